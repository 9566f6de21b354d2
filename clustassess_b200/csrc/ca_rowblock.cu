// ca_rowblock.cu -- the all-pairs ECS engine: row-blocked contingency tables + per-element ECS (sm_100a).
//
// What it replaces: the U(U-1)/2 pair loop of weighted_element_consistency (R/ECS.R:1473-1490) with one
// disjointECS .Call per pair (src/calculate_ecs.cpp:22-46: contingency table, row/col sums, ratio table,
// per-element gather) and the pair loop of element_sim_matrix_flat_disjoint (R/ECS.R:953-959).
//
// Design (B200-first, not a port).  For a partition i the cells are stably sorted by their cluster in i
// (perm_i, ca_pack.cu) so a thread block owns a contiguous run of cells that covers COMPLETE clusters
// of i -- i.e. complete ROWS of every contingency table T_ij.  The rows of one block are small enough
// (<= ~100 KB as 16-bit counters) to live in shared memory, so for each partner partition j the block
//   (1) zeroes its rows and stages S_j (cluster sizes of j) in shared memory,
//   (2) histograms the partner labels of its cells into its rows (shared-memory atomics,
//       thread-local run compression + warp aggregation, so correlated labelings cost ~1 atomic/warp),
//   (3) gathers T[a][b] / max(S_i[a], S_j[b]) per cell -- one IEEE fp64 division, as
//       calculate_ecs.cpp:35 -- and accumulates w_j * ECS into a per-cell register,
// and only after ALL partners writes (w_i/norm) * acc to ecc with one fp64 atomic per cell.
// Labels are read from a cell-major uint16 matrix lab16[n][mpad]: one 16-byte load per cell fetches the
// labels of 8 partner partitions, so HBM traffic is ~2 B per (pair, cell) instead of the 16 B of a
// pair-at-a-time stream, and no table ever leaves the SM.
//
// Work item = (partition i, a set of complete clusters of i).  "Resident" items (<= RB_CAP positions)
// keep cell ids, labels and accumulators in registers; clusters larger than RB_CAP are "streaming"
// items (one table row, cells re-read per phase, ecc accumulated per tile).  Up to JT partners are
// processed per phase (JT tables side by side in shared memory) to amortise the block barriers.
#include "ca_internal.h"

namespace ca {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int CPT = RB_CPT;
static constexpr int THREADS = RB_THREADS;
static constexpr int SMEM_HDR = 128;  // bytes reserved in front of sizes/tables (sim accumulators)

template <bool WIDE>
__device__ __forceinline__ void tbl_add(uint32_t* T, int key, int v) {
    if (WIDE)
        atomicAdd(T + key, (uint32_t)v);
    else  // two 16-bit counters per word; a block never counts past 65535 (planner guarantees it)
        atomicAdd(T + (key >> 1), (uint32_t)v << ((key & 1) << 4));
}
template <bool WIDE>
__device__ __forceinline__ uint32_t tbl_get(const uint32_t* T, int key) {
    if (WIDE) return T[key];
    return (uint32_t) reinterpret_cast<const uint16_t*>(T)[key];
}

__device__ __forceinline__ int lab_of(const uint4& v, int jj) {
    const uint32_t w = (jj >> 1) == 0 ? v.x : (jj >> 1) == 1 ? v.y : (jj >> 1) == 2 ? v.z : v.w;
    return (jj & 1) ? (int)(w >> 16) : (int)(w & 0xffffu);
}

// (2) histogram of one partner partition for this thread's CPT cells.  key = row*kp + b, -1 = padding.
template <bool WIDE>
__device__ __forceinline__ void hist_step(uint32_t* T, const int (&key)[CPT], int lane) {
    const int k0 = key[0];
    int cnt0 = 0;
#pragma unroll
    for (int c = 0; c < CPT; c++) cnt0 += (key[c] == k0) ? 1 : 0;
    bool pending = k0 >= 0;
    // warp aggregation of the threads' leading keys: up to two rounds of "first pending lane broadcasts
    // its key, matching lanes hand it their counts" (covers a warp that straddles two clusters / runs)
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const unsigned pm = __ballot_sync(FULL, pending);
        if (pm == 0) break;
        const int leader = __ffs(pm) - 1;
        const int kl = __shfl_sync(FULL, k0, leader);
        const bool mt = pending && (k0 == kl);
        const int tot = __reduce_add_sync(FULL, mt ? cnt0 : 0);
        if (lane == leader) tbl_add<WIDE>(T, kl, tot);
        pending = pending && !mt;
    }
    if (pending) tbl_add<WIDE>(T, k0, cnt0);
#pragma unroll
    for (int c = 1; c < CPT; c++)
        if (key[c] >= 0 && key[c] != k0) tbl_add<WIDE>(T, key[c], 1);
}

// (3) per-element ECS of one partner partition: e = T[a][b] / max(S_i[a], S_j[b])  (calculate_ecs.cpp:35,39-43)
template <bool WIDE>
__device__ __forceinline__ double gather_step(const uint32_t* T, const uint32_t* szj, const int (&key)[CPT], int rowbase,
                                              uint32_t sa, double wj, double (&acc)[CPT]) {
    const int k0 = key[0];
    double e0 = 0.0, es = 0.0;
    if (k0 >= 0) {
        const uint32_t t = tbl_get<WIDE>(T, k0);
        const uint32_t mx = max(sa, szj[k0 - rowbase]);
        e0 = (double)t / (double)mx;
    }
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        if (key[c] >= 0) {
            double e = e0;
            if (key[c] != k0) {
                const uint32_t t = tbl_get<WIDE>(T, key[c]);
                const uint32_t mx = max(sa, szj[key[c] - rowbase]);
                e = (double)t / (double)mx;
            }
            acc[c] = fma(wj, e, acc[c]);
            es += e;
        }
    }
    return es;
}

template <bool WIDE, bool WANT_ECC, bool WANT_SIM>
__global__ void __launch_bounds__(THREADS, 2) rowblock_kernel(const RowBlockParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Item it = P.items[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31;
    const int m = it.m;
    const int jbeg = it.jbeg, jend = it.jend;
    if (jbeg >= jend) return;
    const int kp = P.kp, nrows = it.nrows;
    constexpr int CW = WIDE ? 4 : 2;
    const int per_j = kp * (4 + CW * nrows);
    int JT = (P.smem_bytes - SMEM_HDR) / per_j;
    JT = JT > RB_JW ? RB_JW : JT;
    double* simacc = reinterpret_cast<double*>(smem);                              // [RB_JW]
    uint32_t* sz = reinterpret_cast<uint32_t*>(smem + SMEM_HDR);                   // [JT][kp]
    uint32_t* tbl = reinterpret_cast<uint32_t*>(smem + SMEM_HDR + (size_t)JT * kp * 4);  // [JT][nrows*kp] counters
    const int tbl_words_j = nrows * kp * CW / 4;
    constexpr int CHUNK = THREADS * CPT;
    const int nsub = (it.npos + CHUNK - 1) / CHUNK;
    const bool resident = nsub == 1;
    const int slot = P.slot_of[m];
    const int32_t* perm = P.perm + (int64_t)slot * P.perm_stride + it.pos0;
    const int mpad = P.mpad;
    const double wscale = P.w[m] * P.inv_norm;

    int cell[CPT];
    uint4 lab[CPT];
    double acc[CPT];
    int rowbase = 0;
    uint32_t sa = (uint32_t)it.sa;
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        cell[c] = -1;
        acc[c] = 0.0;
        lab[c] = make_uint4(0, 0, 0, 0);
    }

    auto load_cells = [&](int s) {
        const int p = s * CHUNK + tid * CPT;
        if (p < it.npos) {
            const int4 v = __ldg(reinterpret_cast<const int4*>(perm + p));
            cell[0] = v.x; cell[1] = v.y; cell[2] = v.z; cell[3] = v.w;
        } else {
            cell[0] = cell[1] = cell[2] = cell[3] = -1;
        }
    };
    auto load_labels = [&](int j0) {
#pragma unroll
        for (int c = 0; c < CPT; c++)
            if (cell[c] >= 0) lab[c] = __ldg(reinterpret_cast<const uint4*>(P.lab16 + (int64_t)cell[c] * mpad + j0));
    };

    if (resident) {
        load_cells(0);
        if (nrows > 1 && cell[0] >= 0) {
            const int a = (int)__ldg(P.lab16 + (int64_t)cell[0] * mpad + m);
            rowbase = (int)__ldg(P.crow + (int64_t)slot * kp + a) * kp;
            sa = (uint32_t)__ldg(P.sizes + (int64_t)m * kp + a);
        }
    }

    for (int j0 = jbeg & ~(RB_JW - 1); j0 < jend; j0 += RB_JW) {
        if (resident) load_labels(j0);
        for (int g0 = 0; g0 < RB_JW; g0 += JT) {
            const int glo = max(g0, jbeg - j0), ghi = min(min(g0 + JT, RB_JW), jend - j0);
            if (glo >= ghi) continue;
            __syncthreads();  // every thread is done gathering from the previous phase's tables
            {   // (1) zero this phase's tables, stage the partners' cluster sizes
                uint4* t4 = reinterpret_cast<uint4*>(tbl);
                const int nq = (ghi - g0) * tbl_words_j / 4;
                for (int i = tid; i < nq; i += THREADS) t4[i] = make_uint4(0, 0, 0, 0);
                const int kq = kp / 4;
                for (int jl = glo - g0; jl < ghi - g0; jl++) {
                    const uint4* src = reinterpret_cast<const uint4*>(P.sizes + (int64_t)(j0 + g0 + jl) * kp);
                    uint4* dst = reinterpret_cast<uint4*>(sz + (size_t)jl * kp);
                    for (int i = tid; i < kq; i += THREADS) dst[i] = __ldg(src + i);
                }
                if (WANT_SIM && tid < RB_JW) simacc[tid] = 0.0;
            }
            __syncthreads();
            // (2) histogram
            for (int s = 0; s < nsub; s++) {
                if (!resident) {
                    load_cells(s);
                    load_labels(j0);
                }
#pragma unroll
                for (int jj = 0; jj < RB_JW; jj++) {
                    if (jj >= glo && jj < ghi) {
                        int key[CPT];
#pragma unroll
                        for (int c = 0; c < CPT; c++) key[c] = cell[c] >= 0 ? rowbase + lab_of(lab[c], jj) : -1;
                        hist_step<WIDE>(tbl + (size_t)(jj - g0) * tbl_words_j, key, lane);
                    }
                }
            }
            __syncthreads();
            // (3) gather + accumulate
            for (int s = 0; s < nsub; s++) {
                if (!resident) {
                    load_cells(s);
                    load_labels(j0);
#pragma unroll
                    for (int c = 0; c < CPT; c++) acc[c] = 0.0;
                }
#pragma unroll
                for (int jj = 0; jj < RB_JW; jj++) {
                    if (jj >= glo && jj < ghi) {
                        int key[CPT];
#pragma unroll
                        for (int c = 0; c < CPT; c++) key[c] = cell[c] >= 0 ? rowbase + lab_of(lab[c], jj) : -1;
                        const double wj = __ldg(P.w + j0 + jj);
                        double es = gather_step<WIDE>(tbl + (size_t)(jj - g0) * tbl_words_j, sz + (size_t)(jj - g0) * kp, key,
                                                      rowbase, sa, wj, acc);
                        if (WANT_SIM) {
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(FULL, es, o);
                            if (lane == 0 && es != 0.0) atomicAdd(simacc + jj, es);
                        }
                    }
                }
                if (!resident && WANT_ECC) {
#pragma unroll
                    for (int c = 0; c < CPT; c++)
                        if (cell[c] >= 0) atomicAdd(P.ecc + cell[c], acc[c] * wscale);
                }
            }
            if (WANT_SIM) {
                __syncthreads();
                if (tid >= glo && tid < ghi && simacc[tid] != 0.0)
                    atomicAdd(P.sim + (int64_t)m * P.m + (j0 + tid), simacc[tid]);
            }
        }
    }
    if (resident && WANT_ECC) {
#pragma unroll
        for (int c = 0; c < CPT; c++)
            if (cell[c] >= 0) atomicAdd(P.ecc + cell[c], acc[c] * wscale);
    }
}

// Shared memory available to one block for sizes + tables, and the number of table rows a resident
// block may own so that at least one partner fits (JT >= 1).  Two blocks per SM when the tables allow.
int rowblock_smem_budget(int32_t kp, int32_t* max_rows_out) {
    Ctx& c = ctx();
    size_t optin = c.smem_optin ? c.smem_optin : 227 * 1024;
    int budget = (int)((optin + 1024) / 2 - 1024 - 512);  // two resident blocks per SM
    int rows = ((budget - SMEM_HDR) / kp - 4) / 2;
    if (rows < 4) {  // large cluster counts: one block per SM
        budget = (int)optin - 512;
        rows = ((budget - SMEM_HDR) / kp - 4) / 2;
    }
    if (rows > RB_MAX_ROWS) rows = RB_MAX_ROWS;
    // do not reserve more shared memory than 8 partners x max_rows rows can use
    long long need = SMEM_HDR + (long long)RB_JW * kp * (4 + 4 * (long long)rows);
    if (need < budget) budget = (int)((need + 15) & ~15LL);
    if (max_rows_out) *max_rows_out = rows;
    return budget;
}

template <bool WIDE, bool E, bool S>
static int launch_one(const RowBlockParams& p, int32_t nitems, cudaStream_t s) {
    auto kern = rowblock_kernel<WIDE, E, S>;
    CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes));
    kern<<<(unsigned)nitems, THREADS, (size_t)p.smem_bytes, s>>>(p);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

int launch_rowblock(const RowBlockParams& p, int32_t nitems, bool wide, cudaStream_t s, int* launches) {
    if (nitems <= 0) return CA_OK;
    const bool e = p.ecc != nullptr, sm = p.sim != nullptr;
    if (launches) (*launches)++;
    if (wide) {
        if (e && sm) return launch_one<true, true, true>(p, nitems, s);
        if (e) return launch_one<true, true, false>(p, nitems, s);
        return launch_one<true, false, true>(p, nitems, s);
    }
    if (e && sm) return launch_one<false, true, true>(p, nitems, s);
    if (e) return launch_one<false, true, false>(p, nitems, s);
    return launch_one<false, false, true>(p, nitems, s);
}

}  // namespace ca
