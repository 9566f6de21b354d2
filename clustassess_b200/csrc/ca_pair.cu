// ca_pair.cu -- direct per-pair kernels (sm_100a) behind the three legacy .Call entry points.
//
//   contingency_smem_kernel   myContTable (src/calculate_ecs.cpp:7-19) for tables that fit in shared
//                             memory: int4-vectorised coalesced label loads, a block-private table,
//                             warp-aggregated shared atomics (__match_any_sync: one atomic per distinct
//                             (a,b) key per warp), non-zero cells flushed with global atomics.
//   contingency_global_kernel the same for ka*kb beyond shared memory: warp-aggregated RED.ADD straight
//                             into the (L2-resident) global table.
//   ecs_gather_kernel         disjointECS (calculate_ecs.cpp:33-43): per element one IEEE fp64 division
//                             T[a][b] / max(S_a, S_b) -- bit-identical to the reference -- plus a
//                             deterministic block-partial sum for mean(ecs).
//   mean_from_table_kernel    disjointECSaverage (calculate_ecs.cpp:62-67): sum T^2 / (n * max).
// Tables are column-major (T[i + j*ka]) exactly like the R matrix the reference returns.
#include "ca_internal.h"

namespace ca {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int PAIR_THREADS = 512;
static constexpr int SMEM_TABLE_MAX = 48 * 1024;  // entries (192 KB of uint32 counters)

__device__ __forceinline__ void add_key(uint32_t* tbl, int64_t key, bool valid, bool shared_space) {
    // warp-aggregated increment: lanes holding the same key elect their first lane to add the group size
    unsigned long long k = valid ? (unsigned long long)key : (0xffffffffffffff00ull | (threadIdx.x & 31));
    unsigned grp = __match_any_sync(FULL, k);
    if (valid && (int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(tbl + key, (uint32_t)__popc(grp));
    (void)shared_space;
}

template <bool SMEM>
__global__ void __launch_bounds__(PAIR_THREADS) contingency_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n,
                                                                   int32_t min_a, int32_t min_b, int32_t ka, int32_t kb,
                                                                   uint32_t* __restrict__ table, uint32_t* __restrict__ sa,
                                                                   uint32_t* __restrict__ sb, int32_t* __restrict__ err) {
    extern __shared__ uint32_t sh[];
    const int64_t cells = (int64_t)ka * kb;
    uint32_t* tbl = SMEM ? sh : table;
    if (SMEM) {
        for (int64_t i = threadIdx.x; i < cells; i += blockDim.x) sh[i] = 0;
        __syncthreads();
    }
    // int4-vectorised main loop over the 16-byte aligned middle, scalar head/tail
    const int64_t n4 = n >> 2;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t tail_from = 0;
    if (aligned) {
        const int4* a4 = reinterpret_cast<const int4*>(a);
        const int4* b4 = reinterpret_cast<const int4*>(b);
        // every lane runs the same trip count so that the warp-collective in add_key stays converged
        const int64_t iters = (n4 + stride - 1) / stride;
        for (int64_t it = 0; it < iters; it++) {
            int64_t i = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
            const bool in = i < n4;
            int4 va = in ? __ldg(a4 + i) : make_int4(0, 0, 0, 0);
            int4 vb = in ? __ldg(b4 + i) : make_int4(0, 0, 0, 0);
            const int av[4] = {va.x, va.y, va.z, va.w};
            const int bv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int64_t ia = (int64_t)av[q] - min_a, ib = (int64_t)bv[q] - min_b;
                bool ok = in && ia >= 0 && ia < ka && ib >= 0 && ib < kb;
                if (in && !ok) *err = 1;
                add_key(tbl, ia + ib * ka, ok, SMEM);
            }
        }
        tail_from = n4 << 2;
    }
    {
        const int64_t rem = n - tail_from;
        const int64_t iters = (rem + stride - 1) / stride;
        for (int64_t it = 0; it < iters; it++) {
            int64_t i = tail_from + it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
            const bool in = i < n;
            int64_t ia = in ? (int64_t)__ldg(a + i) - min_a : 0, ib = in ? (int64_t)__ldg(b + i) - min_b : 0;
            bool ok = in && ia >= 0 && ia < ka && ib >= 0 && ib < kb;
            if (in && !ok) *err = 1;
            add_key(tbl, ia + ib * ka, ok, SMEM);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int64_t i = threadIdx.x; i < cells; i += blockDim.x) {
            uint32_t c = sh[i];
            if (c) {
                atomicAdd(table + i, c);
                if (sa) atomicAdd(sa + (i % ka), c);
                if (sb) atomicAdd(sb + (i / ka), c);
            }
        }
    }
}

// row / column sums of a global table (cluster sizes, calculate_ecs.cpp:31) for the global-table path
__global__ void table_sums_kernel(const uint32_t* __restrict__ table, int32_t ka, int32_t kb, uint32_t* __restrict__ sa,
                                  uint32_t* __restrict__ sb) {
    const int64_t cells = (int64_t)ka * kb;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t c = table[i];
        if (c) {
            atomicAdd(sa + (i % ka), c);
            atomicAdd(sb + (i / ka), c);
        }
    }
}

int pair_contingency(const int32_t* d_a, const int32_t* d_b, int64_t n, int32_t min_a, int32_t min_b, int32_t ka, int32_t kb,
                     uint32_t* d_table, uint32_t* d_sa, uint32_t* d_sb, int32_t* d_err, cudaStream_t s) {
    const int64_t cells = (int64_t)ka * kb;
    CA_CUDA(cudaMemsetAsync(d_table, 0, (size_t)cells * 4, s));
    CA_CUDA(cudaMemsetAsync(d_sa, 0, (size_t)ka * 4, s));
    CA_CUDA(cudaMemsetAsync(d_sb, 0, (size_t)kb * 4, s));
    CA_CUDA(cudaMemsetAsync(d_err, 0, 4, s));
    if (n == 0) return CA_OK;
    Ctx& c = ctx();
    int64_t nb = (n + (int64_t)PAIR_THREADS * 16 - 1) / ((int64_t)PAIR_THREADS * 16);
    if (nb < 1) nb = 1;
    if (cells <= SMEM_TABLE_MAX) {
        if (nb > c.sm_count) nb = c.sm_count;  // one private table per SM: flush cost stays bounded
        size_t smem = (size_t)cells * 4;
        if (smem > 48 * 1024)
            CA_CUDA(cudaFuncSetAttribute(contingency_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        contingency_kernel<true><<<(unsigned)nb, PAIR_THREADS, smem, s>>>(d_a, d_b, n, min_a, min_b, ka, kb, d_table, d_sa, d_sb, d_err);
    } else {
        if (nb > (int64_t)c.sm_count * 4) nb = (int64_t)c.sm_count * 4;
        contingency_kernel<false><<<(unsigned)nb, PAIR_THREADS, 0, s>>>(d_a, d_b, n, min_a, min_b, ka, kb, d_table, nullptr, nullptr, d_err);
        int64_t tb = (cells + 255) / 256;
        if (tb > (int64_t)c.sm_count * 8) tb = (int64_t)c.sm_count * 8;
        table_sums_kernel<<<(unsigned)tb, 256, 0, s>>>(d_table, ka, kb, d_sa, d_sb);
        ca::ctx().launches++;
    }
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (wid == 0) {
        t = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
    }
    return t;  // valid in thread 0
}

__global__ void __launch_bounds__(PAIR_THREADS) ecs_gather_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n,
                                                                  int32_t min_a, int32_t min_b, int32_t ka,
                                                                  const uint32_t* __restrict__ table, const uint32_t* __restrict__ sa,
                                                                  const uint32_t* __restrict__ sb, double* __restrict__ ecs,
                                                                  double* __restrict__ partial) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t ia = (int64_t)__ldg(a + i) - min_a, ib = (int64_t)__ldg(b + i) - min_b;
        uint32_t t = __ldg(table + ia + ib * ka);
        uint32_t mx = max(__ldg(sa + ia), __ldg(sb + ib));
        double e = (double)t / (double)mx;  // calculate_ecs.cpp:35
        if (ecs) ecs[i] = e;
        acc += e;
    }
    double t = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

int pair_ecs_gather(const int32_t* d_a, const int32_t* d_b, int64_t n, int32_t min_a, int32_t min_b, int32_t ka,
                    const uint32_t* d_table, const uint32_t* d_sa, const uint32_t* d_sb, double* d_ecs, double* d_partial,
                    int32_t* nblocks_out, cudaStream_t s) {
    Ctx& c = ctx();
    int64_t nb = (n + (int64_t)PAIR_THREADS * 4 - 1) / ((int64_t)PAIR_THREADS * 4);
    if (nb < 1) nb = 1;
    if (nb > (int64_t)c.sm_count * 4) nb = (int64_t)c.sm_count * 4;
    ecs_gather_kernel<<<(unsigned)nb, PAIR_THREADS, 0, s>>>(d_a, d_b, n, min_a, min_b, ka, d_table, d_sa, d_sb, d_ecs, d_partial);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    *nblocks_out = (int32_t)nb;
    return CA_OK;
}

__global__ void __launch_bounds__(PAIR_THREADS) mean_from_table_kernel(const uint32_t* __restrict__ table, const uint32_t* __restrict__ sa,
                                                                       const uint32_t* __restrict__ sb, int32_t ka, int32_t kb, double n,
                                                                       double* __restrict__ partial) {
    __shared__ double red[32];
    const int64_t cells = (int64_t)ka * kb;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t t = table[i];
        if (t) {
            uint32_t mx = max(sa[i % ka], sb[i / ka]);
            double d = (double)t;
            acc += d * d / (n * (double)mx);  // calculate_ecs.cpp:65 with the product formed in fp64
        }
    }
    double t = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

int pair_mean_from_table(const uint32_t* d_table, const uint32_t* d_sa, const uint32_t* d_sb, int32_t ka, int32_t kb, int64_t n,
                         double* d_partial, int32_t* nblocks_out, cudaStream_t s) {
    Ctx& c = ctx();
    const int64_t cells = (int64_t)ka * kb;
    int64_t nb = (cells + (int64_t)PAIR_THREADS * 4 - 1) / ((int64_t)PAIR_THREADS * 4);
    if (nb < 1) nb = 1;
    if (nb > (int64_t)c.sm_count * 2) nb = (int64_t)c.sm_count * 2;
    mean_from_table_kernel<<<(unsigned)nb, PAIR_THREADS, 0, s>>>(d_table, d_sa, d_sb, ka, kb, (double)n, d_partial);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    *nblocks_out = (int32_t)nb;
    return CA_OK;
}

// fixed-order final reduction of block partials (deterministic): out = scale * sum(partial)
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int32_t nblocks, double scale, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc += partial[i];
    double t = block_sum(acc, red);
    if (threadIdx.x == 0) *out = t * scale;
}

int reduce_partials(const double* d_partial, int32_t nblocks, double scale, double* d_out, cudaStream_t s) {
    reduce_partials_kernel<<<1, 256, 0, s>>>(d_partial, nblocks, scale, d_out);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

// ------------------------------------------------------------------------------------------------
// finish of the batched path: ecc += c0; sim: upper-triangle sums -> means, mirrored, unit diagonal
// (R/ECS.R:1475,1491 scalar terms; R/ECS.R:1479-1482,1488,1497 matrix fill)
// ------------------------------------------------------------------------------------------------
// The wave kernel leaves 64-bit fixed-point sums (ca_internal.h): both kernels convert in place.
__global__ void finish_ecc_kernel(double* ecc, int64_t n, double c0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ecc[i] = __ll2double_rn(reinterpret_cast<const long long*>(ecc)[i]) * (1.0 / ECC_FIX_SCALE) + c0;
}
__global__ void finish_sim_kernel(double* sim, int32_t m, double inv_scale_n) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)m * m) return;
    int i = (int)(idx / m), j = (int)(idx % m);
    if (i == j) {
        sim[idx] = 1.0;
    } else if (i < j) {
        double v = __ll2double_rn(reinterpret_cast<const long long*>(sim)[idx]) * inv_scale_n;
        sim[idx] = v;
        sim[(int64_t)j * m + i] = v;
    }
}

int launch_finish(double* ecc, int64_t n, double c0, double* sim, int32_t m, cudaStream_t s) {
    if (ecc && n > 0) {
        finish_ecc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ecc, n, c0);
        ca::ctx().launches++;
    }
    if (sim) {
        ca::ctx().launches++;
        int64_t t = (int64_t)m * m;
        finish_sim_kernel<<<(unsigned)((t + 255) / 256), 256, 0, s>>>(sim, m, n > 0 ? 1.0 / (sim_fix_scale(n) * (double)n) : 0.0);
    }
    CA_CUDA(cudaGetLastError());
    return CA_OK;
}


// ------------------------------------------------------------------------------------------------
// Any cluster count: one pair through a HASHED contingency table in global memory.
// The batched engine (ca_tile.cu) needs a table row per shared-memory buffer (K <= 24 799) and uint16 labels; the dense
// per-pair table above needs K1 x K2 cells.  Beyond both -- the reference's myContTable (src/calculate_ecs.cpp:7-19) only
// needs the memory -- a pair's table is kept as an open-addressing hash of its non-empty cells (at most N of them):
//   key = (dense rank in a) << 32 | (dense rank in b), 2 N slots rounded up to a power of two, linear probing.
// Two passes like disjointECS (calculate_ecs.cpp:26,39-43): count, then ECS[n] = T[a][b] / max(S_a, S_b) (one IEEE division),
// added to the per-cell fixed-point sums (scaled by w_i w_j / norm) and, summed over the cells, to the pair's matrix entry.
// ------------------------------------------------------------------------------------------------
static constexpr unsigned long long HASH_EMPTY = ~0ull;
__device__ __forceinline__ uint64_t hash_slot(uint64_t key, uint64_t mask) {
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdull;
    key ^= key >> 33;
    key *= 0xc4ceb9fe1a85ec53ull;
    key ^= key >> 33;
    return key & mask;
}
__device__ __forceinline__ uint32_t rank_of(const int32_t* x, int64_t i, int mn, const uint32_t* rk) {
    const uint32_t v = (uint32_t)(__ldg(x + i) - mn);
    return rk ? __ldg(rk + v) : v;
}
__global__ void __launch_bounds__(256) hash_count_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n,
                                                         const int32_t* __restrict__ min_a, const int32_t* __restrict__ min_b,
                                                         const uint32_t* __restrict__ rk_a, const uint32_t* __restrict__ rk_b,
                                                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ counts, uint64_t mask) {
    const int mna = *min_a, mnb = *min_b;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = ((unsigned long long)rank_of(a, i, mna, rk_a) << 32) | rank_of(b, i, mnb, rk_b);
        uint64_t h = hash_slot(key, mask);
        for (;;) {
            const unsigned long long seen = atomicCAS(keys + h, HASH_EMPTY, key);
            if (seen == HASH_EMPTY || seen == key) break;
            h = (h + 1) & mask;
        }
        atomicAdd(counts + h, 1u);
    }
}
__global__ void __launch_bounds__(256) hash_gather_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n,
                                                          const int32_t* __restrict__ min_a, const int32_t* __restrict__ min_b,
                                                          const uint32_t* __restrict__ rk_a, const uint32_t* __restrict__ rk_b,
                                                          const int32_t* __restrict__ sizes_a, const int32_t* __restrict__ sizes_b,
                                                          const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ counts,
                                                          uint64_t mask, double scale, unsigned long long* __restrict__ ecc,
                                                          unsigned long long* __restrict__ sim_slot, double sim_magic) {
    const int mna = *min_a, mnb = *min_b;
    unsigned long long local = 0ull;  // every term rounded to the fixed-point grid before it is added (see ca_tile.cu, sim_fix)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t ca = rank_of(a, i, mna, rk_a), cb = rank_of(b, i, mnb, rk_b);
        const unsigned long long key = ((unsigned long long)ca << 32) | cb;
        uint64_t h = hash_slot(key, mask);
        while (keys[h] != key) h = (h + 1) & mask;
        const double ecs = (double)counts[h] / (double)max(__ldg(sizes_a + ca), __ldg(sizes_b + cb));  // calculate_ecs.cpp:35
        if (ecc) atomicAdd(ecc + i, (unsigned long long)__double2ll_rn(ecs * scale * ECC_FIX_SCALE));
        local += (unsigned long long)(__double_as_longlong(ecs + sim_magic) - __double_as_longlong(sim_magic));
    }
    if (sim_slot) {  // integer sums: the same bits in every run
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL, local, o);
        if ((threadIdx.x & 31) == 0 && local != 0ull) atomicAdd(sim_slot, local);
    }
}

int pair_hashed_accumulate(const int32_t* d_a, const int32_t* d_b, int64_t n, const int32_t* d_min_a, const int32_t* d_min_b,
                           const uint32_t* rk_a, const uint32_t* rk_b, const int32_t* sizes_a, const int32_t* sizes_b, double scale,
                           unsigned long long* ecc, unsigned long long* sim_slot, double sim_magic, cudaStream_t s) {
    uint64_t cap = 1024;
    while (cap < 2 * (uint64_t)n) cap <<= 1;
    DevBuf& hb = scratch(SCR_HASH);
    CA_TRY(hb.ensure((size_t)cap * 12));
    unsigned long long* keys = hb.as<unsigned long long>();
    uint32_t* counts = reinterpret_cast<uint32_t*>(keys + cap);
    CA_CUDA(cudaMemsetAsync(keys, 0xff, (size_t)cap * 8, s));
    CA_CUDA(cudaMemsetAsync(counts, 0, (size_t)cap * 4, s));
    const unsigned nb = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 8));
    hash_count_kernel<<<nb, 256, 0, s>>>(d_a, d_b, n, d_min_a, d_min_b, rk_a, rk_b, keys, counts, cap - 1);
    hash_gather_kernel<<<nb, 256, 0, s>>>(d_a, d_b, n, d_min_a, d_min_b, rk_a, rk_b, sizes_a, sizes_b, keys, counts, cap - 1, scale, ecc, sim_slot,
                                          sim_magic);
    CA_CUDA(cudaGetLastError());
    ca::ctx().launches += 2;
    return CA_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU jobs without NCCL: the primary device adds its peers' fixed-point partial sums, read through peer memory
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) peer_sum_kernel(unsigned long long* __restrict__ dst, const PeerSrc src, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long v = dst[i];
        for (int q = 0; q < src.n; q++) v += src.p[q][i];
        dst[i] = v;
    }
}
int launch_peer_sum(unsigned long long* dst, const PeerSrc& src, int64_t n, cudaStream_t s) {
    if (n <= 0 || src.n <= 0) return CA_OK;
    peer_sum_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 4)), 256, 0, s>>>(dst, src, n);
    CA_CUDA(cudaGetLastError());
    ca::ctx().launches++;
    return CA_OK;
}

}  // namespace ca
