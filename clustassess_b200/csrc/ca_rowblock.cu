// ca_rowblock.cu -- the all-pairs ECS engine: row-blocked contingency tables + per-element ECS (sm_100a).
//
// What it replaces: the U(U-1)/2 pair loop of weighted_element_consistency (R/ECS.R:1473-1490) with one
// disjointECS .Call per pair (src/calculate_ecs.cpp:22-46: contingency table, row/col sums, ratio table,
// per-element gather) and the pair loop of element_sim_matrix_flat_disjoint (R/ECS.R:953-959).
//
// Design (B200-first, not a port).  For a partition i the cells are stably sorted by their cluster in i
// (perm_i, ca_pack.cu) so a thread block owns a contiguous run of cells that covers COMPLETE clusters
// of i -- i.e. complete ROWS of every contingency table T_ij.  The rows of one block are small enough
// (16-bit counters) to live in shared memory, so for each partner partition j the block
//   (1) zeroes its rows and stages S_j (cluster sizes of j) in shared memory,
//   (2) histograms the partner labels of its cells into its rows: a thread owns 4 consecutive cells,
//       counts how many share the label of its first cell, threads with the same leading key are
//       aggregated across the warp (ballot / shfl / redux, two rounds), so a run of equal labels costs
//       ONE shared atomic per warp; only the cells that differ ("exceptions") add individually,
//   (3) gathers ECS = T[a][b] / max(S_i[a], S_j[b])  (calculate_ecs.cpp:35) once per thread for the
//       leading key; the exceptions of the whole warp are compacted into a dense list (ballot + popc) and
//       evaluated one per lane; everything is accumulated in per-cell registers,
// and only after ALL partners writes (w_i/norm) * acc to ecc with one fp64 atomic per cell.
// Labels come from a cell-major uint16 matrix lab16[n][mpad]: one 16-byte load per cell fetches the labels
// of 8 partner partitions (kept in a per-thread shared-memory scratch tile so that the partner loop is a
// rolled loop with running addresses), so HBM traffic is ~2 B per (pair, cell) instead of the 16 B of a
// pair-at-a-time stream, and no table ever leaves the SM.
//
// Work item = (partition i, a set of complete clusters of i).  "Resident" items (<= RB_CAP positions)
// keep their accumulators in registers for the whole job; clusters larger than RB_CAP are "streaming"
// items (one table row, cells processed in sub-chunks, ecc accumulated per phase).  Up to JT partners are
// processed per phase (JT tables side by side in shared memory) to amortise the block barriers.
//
// The division T / max(Sa, Sb) is evaluated as T * r with r = 1/max from rcp.approx.f64 refined by two
// Newton steps (relative error ~1e-16): the batched path only has to meet 1e-9 absolute; the per-pair entry
// points (ca_pair.cu) keep the single IEEE division and are bit-identical to the reference.
#include "ca_internal.h"

namespace ca {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int CPT = RB_CPT;
static constexpr int THREADS = RB_THREADS;
static constexpr int NWARPS = THREADS / 32;
static_assert(CPT == 4, "the packed-label code below is written for 4 cells per thread");

// shared-memory map (byte offsets from the dynamic shared base)
static constexpr int STAGE_REC = 64;                          // exception records per warp and pass
static constexpr int SM_SIMACC = 0;                           // RB_JW doubles
static constexpr int SM_MBAR = 64;                            // two mbarriers: [0] tables zeroed, [1] sizes staged
static constexpr int SM_STAGE = 128;                          // per warp STAGE_REC x 8 B records / results
static constexpr int SM_LAB = SM_STAGE + NWARPS * STAGE_REC * 8;   // [RB_JW][THREADS] x 8 B: 4 packed labels per thread
static constexpr int SM_DYN = SM_LAB + RB_JW * THREADS * 8;   // [JT][kp] uint32 sizes of the partners, then [JT][nrows*kp] counters

// ---- 32-bit shared-space accessors ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void lds_v2(uint32_t a, uint32_t& x, uint32_t& y) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void red_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add_f64(uint32_t a, double v) { asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier -------------------------------
// One thread stages the partners' cluster sizes and zero-fills the tables (copying from a zero buffer in global
// memory) without spending LSU instructions or registers; the block waits on the barrier's phase parity.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();  // a lost copy must fail loudly, never hang the GPU
    }
}

// counter of table entry `key` (table base tb, bytes): two 16-bit counters per word unless WIDE
template <bool WIDE>
__device__ __forceinline__ void tbl_add(uint32_t tb, uint32_t key, uint32_t v) {
    if (WIDE)
        red_add(tb + key * 4u, v);
    else  // a block never counts past 65535 (the planner guarantees it), so halves cannot carry
        red_add(tb + ((key * 2u) & ~3u), v << ((key & 1u) << 4));
}
template <bool WIDE>
__device__ __forceinline__ uint32_t tbl_get(uint32_t tb, uint32_t key) {
    return WIDE ? lds_u32(tb + key * 4u) : lds_u16(tb + key * 2u);
}

// T / max(sa, sb): reciprocal seed + two Newton steps, then one multiply
__device__ __forceinline__ double ecs_of(uint32_t count, uint32_t sa, uint32_t sb) {
    const double mx = (double)max(sa, sb);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(mx));
    r = fma(fma(-mx, r, 1.0), r, r);
    r = fma(fma(-mx, r, 1.0), r, r);
    return (double)count * r;
}

struct Acc {           // per-thread accumulators: acc(cell c) = base + corr[c]
    double base;       // sum_j w_j * ECS of the thread's leading key
    double corr[CPT];  // corrections for the partners where cell c differed from the leading key
};

// (2) histogram of one partner for this thread's 4 cells (p01, p23 = packed labels; padding cells mirror cell 0).
// Cost model measured on B200 (tools/microbench.cu, profiles/): a shared ATOMS instruction costs ~11 SM-cycles
// of the LSU pipe however few lanes are active, and lanes that hit the SAME word serialise into one wavefront
// each.  So a warp issues ONE atomic instruction per partner whose lanes hold DISTINCT keys:
//   - the threads' leading keys are aggregated (ballot / shfl / redux, two rounds: a warp rarely spans more
//     than two runs) into one (key, count) record per run,
//   - cells that differ from their thread's leading key become (key, 1) records,
//   - all records are compacted into a dense list (ballot + popc -> per-warp staging buffer) and applied 32
//     per atomic instruction.
template <bool WIDE>
__device__ __forceinline__ void hist_step(uint32_t tb, uint32_t stage, uint32_t p01, uint32_t p23, uint32_t rowbase, int nvalid,
                                          int lane) {
    const uint32_t kk = __byte_perm(p01, 0, 0x1010);  // leading label in both halves
    const uint32_t d01 = p01 ^ kk, d23 = p23 ^ kk;    // non-zero halves = cells that differ from it
    const bool x1 = (d01 >> 16) != 0, x2 = (d23 & 0xffffu) != 0, x3 = (d23 >> 16) != 0;
    const uint32_t k0 = rowbase + (p01 & 0xffffu);
    const int cnt0 = nvalid - (int)x1 - (int)x2 - (int)x3;
    bool pending = cnt0 > 0;
    uint32_t rec0 = k0 | ((uint32_t)cnt0 << 16);  // record = key | count << 16 (key < 65536: the table is in smem)
    bool x0 = pending;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const unsigned pm = __ballot_sync(FULL, pending);
        if (pm == 0) break;
        const int leader = __ffs(pm) - 1;
        const uint32_t kl = __shfl_sync(FULL, k0, leader);
        const bool mt = pending && (k0 == kl);
        const int tot = __reduce_add_sync(FULL, mt ? cnt0 : 0);
        if (mt) {
            x0 = lane == leader;  // only the run's first thread keeps a record
            rec0 = k0 | ((uint32_t)tot << 16);
        }
        pending = pending && !mt;
    }
    const unsigned m0 = __ballot_sync(FULL, x0), m1 = __ballot_sync(FULL, x1), m2 = __ballot_sync(FULL, x2),
                   m3 = __ballot_sync(FULL, x3);
    const unsigned lt = (1u << lane) - 1u;
    const int n0 = __popc(m0), n1 = n0 + __popc(m1), n2 = n1 + __popc(m2), ne = n2 + __popc(m3);  // ne <= 128
    if (x0) sts_u32(stage + (uint32_t)__popc(m0 & lt) * 4u, rec0);
    if (x1) sts_u32(stage + (uint32_t)(n0 + __popc(m1 & lt)) * 4u, (rowbase + (p01 >> 16)) | 0x10000u);
    if (x2) sts_u32(stage + (uint32_t)(n1 + __popc(m2 & lt)) * 4u, (rowbase + (p23 & 0xffffu)) | 0x10000u);
    if (x3) sts_u32(stage + (uint32_t)(n2 + __popc(m3 & lt)) * 4u, (rowbase + (p23 >> 16)) | 0x10000u);
    __syncwarp();
    for (int h = lane; h < ne; h += 32) {
        const uint32_t rec = lds_u32(stage + h * 4u);
        tbl_add<WIDE>(tb, rec & 0xffffu, rec >> 16);
    }
    __syncwarp();
}

// (3) per-element ECS of one partner.  Returns this lane's share of sum_cells ECS (for the ECS matrix).
// Partner cluster sizes come from shared memory (szb, staged per phase) or, with GSZ, straight from the global
// sizes row szg (L2-resident, 4 MB in all): the leading key's size sb0 is then fetched by the caller BEFORE it
// waits for the table, so only the exceptions' lookups sit in the dependent chain.
template <bool WIDE, bool WANT_SIM, bool GSZ = false>
__device__ __forceinline__ double gather_step(uint32_t tb, uint32_t szb, uint32_t stage, uint32_t p01, uint32_t p23,
                                              uint32_t rowbase, uint32_t sa, int nvalid, double wj, int lane, Acc& acc,
                                              const int32_t* __restrict__ szg = nullptr, uint32_t sb0 = 0) {
    const uint32_t kk = __byte_perm(p01, 0, 0x1010);
    const uint32_t d01 = p01 ^ kk, d23 = p23 ^ kk;
    const bool x1 = (d01 >> 16) != 0, x2 = (d23 & 0xffffu) != 0, x3 = (d23 >> 16) != 0;
    const uint32_t b0 = p01 & 0xffffu;
    double we0 = 0.0, es = 0.0;
    if (nvalid > 0) {
        const double e0 = ecs_of(tbl_get<WIDE>(tb, rowbase + b0), sa, GSZ ? sb0 : lds_u32(szb + b0 * 4u));
        we0 = e0 * wj;
        acc.base += we0;
        if (WANT_SIM) es = e0 * (double)(nvalid - (int)x1 - (int)x2 - (int)x3);
    }
    const unsigned m1 = __ballot_sync(FULL, x1), m2 = __ballot_sync(FULL, x2), m3 = __ballot_sync(FULL, x3);
    if ((m1 | m2 | m3) != 0) {
        // compact the warp's exceptions into a dense list: one lane evaluates one exception
        const unsigned lt = (1u << lane) - 1u;
        const int n1 = __popc(m1), n2 = __popc(m2);
        const int i1 = __popc(m1 & lt), i2 = n1 + __popc(m2 & lt), i3 = n1 + n2 + __popc(m3 & lt);
        const int ne = n1 + n2 + __popc(m3);
        for (int q0 = 0; q0 < ne; q0 += STAGE_REC) {  // one pass unless a warp has > STAGE_REC exceptions
            const bool o1 = x1 && (unsigned)(i1 - q0) < (unsigned)STAGE_REC;
            const bool o2 = x2 && (unsigned)(i2 - q0) < (unsigned)STAGE_REC;
            const bool o3 = x3 && (unsigned)(i3 - q0) < (unsigned)STAGE_REC;
            // record = { key | label << 16 , S_i[row] }; key < 65536 because the table lives in shared memory
            if (o1) sts_v2(stage + (i1 - q0) * 8u, (rowbase + (p01 >> 16)) | (p01 & 0xffff0000u), sa);
            if (o2) sts_v2(stage + (i2 - q0) * 8u, (rowbase + (p23 & 0xffffu)) | (p23 << 16), sa);
            if (o3) sts_v2(stage + (i3 - q0) * 8u, (rowbase + (p23 >> 16)) | (p23 & 0xffff0000u), sa);
            __syncwarp();
            const int nq = min(ne - q0, STAGE_REC);
            for (int h = lane; h < nq; h += 32) {
                uint32_t rec, rsa;
                lds_v2(stage + h * 8u, rec, rsa);
                const uint32_t rsb = GSZ ? (uint32_t)__ldg(szg + (rec >> 16)) : lds_u32(szb + (rec >> 16) * 4u);
                const double e = ecs_of(tbl_get<WIDE>(tb, rec & 0xffffu), rsa, rsb);
                if (WANT_SIM) es += e;
                sts_f64(stage + h * 8u, e * wj);
            }
            __syncwarp();
            if (o1) acc.corr[1] += lds_f64(stage + (i1 - q0) * 8u) - we0;
            if (o2) acc.corr[2] += lds_f64(stage + (i2 - q0) * 8u) - we0;
            if (o3) acc.corr[3] += lds_f64(stage + (i3 - q0) * 8u) - we0;
            __syncwarp();
        }
    }
    return es;
}

template <bool WIDE, bool RESIDENT, bool WANT_ECC, bool WANT_SIM>
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
    int JT = (P.smem_bytes - SM_DYN) / per_j;
    JT = JT > RB_JW ? RB_JW : JT;
#ifdef CA_ABLATE
    if (P.dbg >> 8) JT = min(JT, P.dbg >> 8);  // force a smaller partner group
#endif
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t s_stage = sbase + SM_STAGE + (tid >> 5) * (STAGE_REC * 8);
    const uint32_t s_lab = sbase + SM_LAB + tid * 8u;             // + jj * THREADS * 8
    const uint32_t s_sz = sbase + SM_DYN;                         // [JT][kp] uint32
    const uint32_t s_tbl = s_sz + (uint32_t)JT * kp * 4u;         // [JT][nrows*kp] counters
    const uint32_t tbl_bytes_j = (uint32_t)nrows * kp * CW;
    constexpr int CHUNK = THREADS * CPT;
    const int nsub = RESIDENT ? 1 : (it.npos + CHUNK - 1) / CHUNK;
    const int slot = P.slot_of[m];
    const int32_t* perm = P.perm + (int64_t)slot * P.perm_stride + it.pos0;
    const int mpad = P.mpad;
    const double wscale = P.w[m] * P.inv_norm;
#ifdef CA_ABLATE
    const int dbg = P.dbg;
#else
    constexpr int dbg = 0;
#endif

    int4 cell = make_int4(-1, -1, -1, -1);
    int nvalid = 0;
    uint32_t rowbase = 0, sa = (uint32_t)it.sa;
    Acc acc;
    acc.base = 0.0;
#pragma unroll
    for (int c = 0; c < CPT; c++) acc.corr[c] = 0.0;

    auto load_cells = [&](int s) {
        const int p = s * CHUNK + tid * CPT;
        cell = make_int4(-1, -1, -1, -1);
        if (p < it.npos) cell = __ldg(reinterpret_cast<const int4*>(perm + p));
        nvalid = (cell.x >= 0) + (cell.y >= 0) + (cell.z >= 0) + (cell.w >= 0);
    };
    // Fetch this thread's 4 x 8 labels of the tile [j0, j0+8) and park them, partner-major, in its private
    // slice of the shared scratch tile: s_lab + jj*THREADS*8 holds {cell0|cell1<<16, cell2|cell3<<16}.
    // Padding cells mirror cell 0 (never an exception); idle threads store zeros.  No barrier is needed:
    // a thread only ever reads back what it wrote itself.
    auto load_labels = [&](int j0) {
        uint4 l0 = make_uint4(0, 0, 0, 0), l1, l2, l3;
        if (cell.x >= 0) l0 = __ldg(reinterpret_cast<const uint4*>(P.lab16 + (int64_t)cell.x * mpad + j0));
        l1 = l2 = l3 = l0;
        if (cell.y >= 0) l1 = __ldg(reinterpret_cast<const uint4*>(P.lab16 + (int64_t)cell.y * mpad + j0));
        if (cell.z >= 0) l2 = __ldg(reinterpret_cast<const uint4*>(P.lab16 + (int64_t)cell.z * mpad + j0));
        if (cell.w >= 0) l3 = __ldg(reinterpret_cast<const uint4*>(P.lab16 + (int64_t)cell.w * mpad + j0));
        const uint32_t w0[4] = {l0.x, l0.y, l0.z, l0.w}, w1[4] = {l1.x, l1.y, l1.z, l1.w};
        const uint32_t w2[4] = {l2.x, l2.y, l2.z, l2.w}, w3[4] = {l3.x, l3.y, l3.z, l3.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            sts_v2(s_lab + (2 * q) * (THREADS * 8), __byte_perm(w0[q], w1[q], 0x5410), __byte_perm(w2[q], w3[q], 0x5410));
            sts_v2(s_lab + (2 * q + 1) * (THREADS * 8), __byte_perm(w0[q], w1[q], 0x7632), __byte_perm(w2[q], w3[q], 0x7632));
        }
    };
    auto flush = [&]() {
        if (cell.x >= 0) atomicAdd(P.ecc + cell.x, acc.base * wscale);
        if (cell.y >= 0) atomicAdd(P.ecc + cell.y, (acc.base + acc.corr[1]) * wscale);
        if (cell.z >= 0) atomicAdd(P.ecc + cell.z, (acc.base + acc.corr[2]) * wscale);
        if (cell.w >= 0) atomicAdd(P.ecc + cell.w, (acc.base + acc.corr[3]) * wscale);
    };

    if (RESIDENT) {
        load_cells(0);
        if (nrows > 1 && cell.x >= 0) {
            const int a = (int)__ldg(P.lab16 + (int64_t)cell.x * mpad + m);
            rowbase = (uint32_t)__ldg(P.crow + (int64_t)slot * kp + a) * (uint32_t)kp;
            sa = (uint32_t)__ldg(P.sizes + (int64_t)m * kp + a);
        }
    }

    const uint32_t bar_z = sbase + SM_MBAR, bar_s = sbase + SM_MBAR + 8;
    if (tid == 0) {
        mbar_init(bar_z, 1);
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t par = 0;  // phase parity of both barriers
    const int jfirst = jbeg & ~(RB_JW - 1);
    for (int j0 = jfirst; j0 < jend; j0 += RB_JW) {
        if (RESIDENT && !(dbg & 8)) {
            load_labels(j0);
        }
        const int tlo = max(0, jbeg - j0), thi = min(RB_JW, jend - j0);  // partners [tlo, thi) of this tile are wanted
        for (int g0 = tlo; g0 < thi; g0 += JT) {
            const int ng = min(JT, thi - g0);  // partners of this phase: j0+g0 .. j0+g0+ng-1
            __syncthreads();  // every thread is done gathering from the previous phase's tables
            if (!(dbg & 4)) {   // (1) zero this phase's tables and stage the partners' cluster sizes: two bulk copies
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order the block's generic reads before the async writes
                    const uint32_t nz = (uint32_t)ng * tbl_bytes_j, nsz = (uint32_t)ng * kp * 4u;
                    mbar_expect_tx(bar_z, nz);
                    bulk_g2s(s_tbl, P.zeros, nz, bar_z);
                    mbar_expect_tx(bar_s, nsz);
                    bulk_g2s(s_sz, P.sizes + (int64_t)(j0 + g0) * kp, nsz, bar_s);
                }
                if (WANT_SIM && tid < RB_JW) sts_f64(sbase + SM_SIMACC + tid * 8u, 0.0);
                mbar_wait(bar_z, par);
            }
            // (2) histogram
            for (int s = 0; s < ((dbg & 1) ? 0 : nsub); s++) {
                if (!RESIDENT) {
                    load_cells(s);
                    load_labels(j0);
                }
                uint32_t tb = s_tbl, la = s_lab + (uint32_t)g0 * (THREADS * 8);
                for (int jl = 0; jl < ng; jl++, tb += tbl_bytes_j, la += THREADS * 8) {
                    uint32_t p01, p23;
                    lds_v2(la, p01, p23);
                    hist_step<WIDE>(tb, s_stage, p01, p23, rowbase, nvalid, lane);
                }
            }
            __syncthreads();
            if (!(dbg & 4)) mbar_wait(bar_s, par);
            par ^= 1u;
            // (3) gather + accumulate
            for (int s = 0; s < ((dbg & 2) ? 0 : nsub); s++) {
                if (!RESIDENT) {
                    load_cells(s);
                    load_labels(j0);
                    acc.base = 0.0;
#pragma unroll
                    for (int c = 0; c < CPT; c++) acc.corr[c] = 0.0;
                }
                uint32_t tb = s_tbl, szb = s_sz, la = s_lab + (uint32_t)g0 * (THREADS * 8);
                for (int jl = 0; jl < ng; jl++, tb += tbl_bytes_j, szb += (uint32_t)kp * 4u, la += THREADS * 8) {
                    uint32_t p01, p23;
                    lds_v2(la, p01, p23);
                    const double wj = __ldg(P.w + j0 + g0 + jl);
                    double es = gather_step<WIDE, WANT_SIM>(tb, szb, s_stage, p01, p23, rowbase, sa, nvalid, wj, lane, acc);
                    if (WANT_SIM) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(FULL, es, o);
                        if (lane == 0 && es != 0.0) red_add_f64(sbase + SM_SIMACC + jl * 8u, es);
                    }
                }
                if (!RESIDENT && WANT_ECC) flush();
            }
            if (WANT_SIM) {
                __syncthreads();
                if (tid < ng) {
                    const double v = lds_f64(sbase + SM_SIMACC + tid * 8u);
                    if (v != 0.0) atomicAdd(P.sim + (int64_t)m * P.m + (j0 + g0 + tid), v);
                }
            }
        }
    }
    if (RESIDENT && WANT_ECC) flush();
}

// ---------------------------------------------------------------------------------------------------------
// Pipelined resident kernel.  Measured on the full config-5 job, the phase-synchronous kernel above spends
// ~60 % of its time stalled at the two block barriers of every phase (warp skew + staging latency): its time
// is A + B/JT with JT (partners per phase) capped at ~3 by shared memory.  Here the same shared memory is a
// RING of NB >= 3 buffers, one partner each, and all hand-offs are mbarriers, so no warp ever waits for the
// slowest warp of the SAME step:
//     iteration g (every warp):  hist(partner g) -> arrive H[g]        ... table of g is being built
//                                wait H[g-1] (all warps)               ... finished one half-iteration ago
//                                gather(partner g-1) -> arrive G[g-1]
//     the LAST warp to finish gather(g-1) (a shared release counter tells it) refills that buffer at once for
//     partner g-1+NB with a TMA bulk copy from a zero buffer, completing on Z.
// Nobody ever waits for a refill to be ISSUED, every dependency has at least half an iteration of slack, and the
// TMA latency has a whole iteration.  Partner cluster sizes are NOT staged: the leading key's size is fetched
// from the L2-resident sizes matrix before the H wait, the exceptions' sizes by their handler lanes -- which frees
// a third of the ring's shared memory for more buffers.
// ---------------------------------------------------------------------------------------------------------
static constexpr int PM_SIMACC = 0;                            // RB_JW doubles, one per ring buffer
static constexpr int PM_MBAR = 64;                             // 3 x RB_JW mbarriers (Z, S, H per buffer) + RB_JW release counters
static constexpr int PM_STAGE = PM_MBAR + 4 * RB_JW * 8;
static constexpr int PM_LAB = PM_STAGE + NWARPS * STAGE_REC * 8;
static constexpr int PM_DYN = PM_LAB + RB_JW * THREADS * 8;    // ring: NB x nrows*kp 16-bit counters

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

int rowblock_pipe_buffers(int32_t kp, int32_t nrows, int32_t smem_bytes) {
    const int per_j = kp * 2 * nrows;
    const int nb = (smem_bytes - PM_DYN) / per_j;
    return nb > RB_JW ? RB_JW : nb;
}

template <bool WANT_ECC, bool WANT_SIM>
__global__ void __launch_bounds__(THREADS, 2) rowblock_pipe_kernel(const RowBlockParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Item it = P.items[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31;
    const int m = it.m;
    const int jbeg = it.jbeg, G = it.jend - it.jbeg;
    if (G <= 0) return;
    const int kp = P.kp, nrows = it.nrows;
    const uint32_t per_j = (uint32_t)kp * 2u * nrows;  // one table per ring buffer
    int NB = (P.smem_bytes - PM_DYN) / (int)per_j;
    NB = NB > RB_JW ? RB_JW : NB;  // >= 3 guaranteed by the host
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t s_stage = sbase + PM_STAGE + (tid >> 5) * (STAGE_REC * 8);
    const uint32_t s_lab = sbase + PM_LAB + tid * 8u;
    const uint32_t s_ring = sbase + PM_DYN;                       // buffer b: table at s_ring + b*per_j
    const uint32_t bar0 = sbase + PM_MBAR;                        // Z[b] = bar0+b*8, H[b] = +128, release counter = +192
    const int slot = P.slot_of[m];
    const int32_t* perm = P.perm + (int64_t)slot * P.perm_stride + it.pos0;
    const int mpad = P.mpad;
    const double wscale = P.w[m] * P.inv_norm;

    if (tid == 0) {
        for (int b = 0; b < RB_JW; b++) {
            mbar_init(bar0 + b * 8, 1);
            mbar_init(bar0 + 128 + b * 8, NWARPS);
            sts_u32(bar0 + 192 + b * 8, 0u);  // release counter of buffer b
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (WANT_SIM && tid < RB_JW) sts_f64(sbase + PM_SIMACC + tid * 8u, 0.0);

    int4 cell = make_int4(-1, -1, -1, -1);
    {
        const int p = tid * CPT;
        if (p < it.npos) cell = __ldg(reinterpret_cast<const int4*>(perm + p));
    }
    const int nvalid = (cell.x >= 0) + (cell.y >= 0) + (cell.z >= 0) + (cell.w >= 0);
    uint32_t rowbase = 0, sa = (uint32_t)it.sa;
    if (nrows > 1 && cell.x >= 0) {
        const int a = (int)__ldg(P.lab16 + (int64_t)cell.x * mpad + m);
        rowbase = (uint32_t)__ldg(P.crow + (int64_t)slot * kp + a) * (uint32_t)kp;
        sa = (uint32_t)__ldg(P.sizes + (int64_t)m * kp + a);
    }
    Acc acc;
    acc.base = 0.0;
#pragma unroll
    for (int c = 0; c < CPT; c++) acc.corr[c] = 0.0;

    // A cell's 32-byte DRAM sector holds the labels of TWO consecutive 8-partner tiles.  The even tile's load asks
    // L2 to keep the sector (evict_last) so that the odd tile, one tile period later, still finds it; the odd
    // tile's load releases it (evict_first).  Otherwise every sector is fetched from DRAM twice.
    uint64_t pol_keep, pol_drop;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
    auto ldg_hint = [&](const uint16_t* ptr, uint64_t pol) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(ptr), "l"(pol));
        return v;
    };
    auto load_labels = [&](int j0) {  // as in the phase-synchronous kernel: thread-private slice of the scratch tile
        const uint64_t pol = (j0 & RB_JW) ? pol_drop : pol_keep;
        uint4 l0 = make_uint4(0, 0, 0, 0), l1, l2, l3;
        if (cell.x >= 0) l0 = ldg_hint(P.lab16 + (int64_t)cell.x * mpad + j0, pol);
        l1 = l2 = l3 = l0;
        if (cell.y >= 0) l1 = ldg_hint(P.lab16 + (int64_t)cell.y * mpad + j0, pol);
        if (cell.z >= 0) l2 = ldg_hint(P.lab16 + (int64_t)cell.z * mpad + j0, pol);
        if (cell.w >= 0) l3 = ldg_hint(P.lab16 + (int64_t)cell.w * mpad + j0, pol);
        const uint32_t w0[4] = {l0.x, l0.y, l0.z, l0.w}, w1[4] = {l1.x, l1.y, l1.z, l1.w};
        const uint32_t w2[4] = {l2.x, l2.y, l2.z, l2.w}, w3[4] = {l3.x, l3.y, l3.z, l3.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            sts_v2(s_lab + (2 * q) * (THREADS * 8), __byte_perm(w0[q], w1[q], 0x5410), __byte_perm(w2[q], w3[q], 0x5410));
            sts_v2(s_lab + (2 * q + 1) * (THREADS * 8), __byte_perm(w0[q], w1[q], 0x7632), __byte_perm(w2[q], w3[q], 0x7632));
        }
    };
    // refill ring buffer of partner gz (one thread): flush + clear its ECS-matrix accumulator, zero-fill the table
    auto refill = [&](int gz) {
        const int b = gz % NB;
        if (WANT_SIM && gz >= NB) {
            const double v = lds_f64(sbase + PM_SIMACC + b * 8u);
            if (v != 0.0) atomicAdd(P.sim + (int64_t)m * P.m + (jbeg + gz - NB), v);
            sts_f64(sbase + PM_SIMACC + b * 8u, 0.0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the block's generic reads before the async writes
        mbar_expect_tx(bar0 + b * 8, per_j);
        bulk_g2s(s_ring + (uint32_t)b * per_j, P.zeros, per_j, bar0 + b * 8);
    };

    __syncthreads();  // barriers initialised, accumulators cleared
    if (tid == 0)
        for (int gz = 0; gz < NB && gz < G; gz++) refill(gz);

    const int jfirst = jbeg & ~(RB_JW - 1);
    uint32_t q01 = 0, q23 = 0;  // labels of the previous partner (gathered one iteration after their histogram)
#ifdef CA_ABLATE
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = clock64(), tstart = t0;
#define CA_TICK(k) { long long t1_ = clock64(); tk[k] += t1_ - t0; t0 = t1_; }
#else
#define CA_TICK(k)
#endif
    for (int g = 0; g <= G; g++) {
        uint32_t p01 = 0, p23 = 0;
        if (g < G) {
            const int j = jbeg + g, jj = (j - jfirst) & (RB_JW - 1);
            if (jj == 0 || g == 0) load_labels(j & ~(RB_JW - 1));
            const int b = g % NB;
            lds_v2(s_lab + (uint32_t)jj * (THREADS * 8), p01, p23);
            CA_TICK(0)
            mbar_wait(bar0 + b * 8, (uint32_t)((g / NB) & 1));  // table zero-filled
            CA_TICK(1)
            hist_step<false>(s_ring + (uint32_t)b * per_j, s_stage, p01, p23, rowbase, nvalid, lane);
            if (lane == 0) mbar_arrive(bar0 + 128 + b * 8);      // (hist_step ends with __syncwarp)
            CA_TICK(2)
        }
        CA_TICK(3)
        if (g >= 1) {
            const int b = (g - 1) % NB;
            const int32_t* szg = P.sizes + (int64_t)(jbeg + g - 1) * kp;  // partner g-1's cluster sizes (global, L2)
            const uint32_t sb0 = nvalid > 0 ? (uint32_t)__ldg(szg + (q01 & 0xffffu)) : 0u;  // in flight during the wait
            const double wj = __ldg(P.w + jbeg + g - 1);
            mbar_wait(bar0 + 128 + b * 8, (uint32_t)(((g - 1) / NB) & 1));  // every warp has added partner g-1 to the table
            CA_TICK(4)
            double es = gather_step<false, WANT_SIM, true>(s_ring + (uint32_t)b * per_j, 0u, s_stage, q01, q23, rowbase, sa, nvalid, wj,
                                                           lane, acc, szg, sb0);
            if (WANT_SIM) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(FULL, es, o);
                if (lane == 0 && es != 0.0) red_add_f64(sbase + PM_SIMACC + b * 8u, es);
            }
            __syncwarp();
            if (lane == 0) {  // release the buffer; the last warp to do so refills it for partner g-1+NB right away
                uint32_t old;
                asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(bar0 + 192 + b * 8) : "memory");
                if (old == NWARPS - 1) {
                    sts_u32(bar0 + 192 + b * 8, 0u);
                    if (g - 1 + NB < G) refill(g - 1 + NB);
                }
            }
            CA_TICK(6)
        }
        q01 = p01;
        q23 = p23;
    }
#ifdef CA_ABLATE
    if (P.timeline && lane == 0 && blockIdx.x < 256) {
        long long* o = P.timeline + ((size_t)blockIdx.x * NWARPS + (tid >> 5)) * 10;
        for (int k = 0; k < 7; k++) o[k] = tk[k];
        o[7] = clock64() - tstart;
        o[8] = G;
        o[9] = NB;
    }
#endif
    if (WANT_SIM) {  // the last min(NB, G) partners still sit in their accumulators
        __syncthreads();
        if (tid < NB && tid < G) {
            const int gl = G - 1 - ((G - 1 - tid) % NB + NB) % NB;  // last partner that used buffer tid
            const double v = lds_f64(sbase + PM_SIMACC + tid * 8u);
            if (gl >= 0 && v != 0.0) atomicAdd(P.sim + (int64_t)m * P.m + (jbeg + gl), v);
        }
    }
    if (WANT_ECC) {
        if (cell.x >= 0) atomicAdd(P.ecc + cell.x, acc.base * wscale);
        if (cell.y >= 0) atomicAdd(P.ecc + cell.y, (acc.base + acc.corr[1]) * wscale);
        if (cell.z >= 0) atomicAdd(P.ecc + cell.z, (acc.base + acc.corr[2]) * wscale);
        if (cell.w >= 0) atomicAdd(P.ecc + cell.w, (acc.base + acc.corr[3]) * wscale);
    }
}

// Shared memory available to one block, and the number of table rows a resident block may own so that
// at least one partner fits (JT >= 1).  Two blocks per SM when the tables allow it.
int rowblock_smem_budget(int32_t kp, int32_t* max_rows_out) {
    Ctx& c = ctx();
    size_t optin = c.smem_optin ? c.smem_optin : 227 * 1024;
    int budget = (int)((optin + 1024) / 2 - 1024 - 512);  // two resident blocks per SM
    int rows = ((budget - SM_DYN) / kp - 4) / 2;
    if (rows < 4) {  // large cluster counts: one block per SM
        budget = (int)optin - 512;
        rows = ((budget - SM_DYN) / kp - 4) / 2;
    }
    if (rows > RB_MAX_ROWS) rows = RB_MAX_ROWS;
    if (rows > 65536 / kp) rows = 65536 / kp > 0 ? 65536 / kp : 1;  // exception records carry 16-bit table keys
    // do not reserve more shared memory than 8 partners x max_rows rows can use
    long long need = SM_DYN + (long long)RB_JW * kp * (4 + 4 * (long long)(rows > 0 ? rows : 1));
    if (need < budget) budget = (int)((need + 15) & ~15LL);
    if (max_rows_out) *max_rows_out = rows;
    return budget;
}

template <bool WIDE, bool R, bool E, bool S>
static int launch_one(const RowBlockParams& p, int32_t nitems, cudaStream_t s) {
    auto kern = rowblock_kernel<WIDE, R, E, S>;
    CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes));
    kern<<<(unsigned)nitems, THREADS, (size_t)p.smem_bytes, s>>>(p);
    CA_CUDA(cudaGetLastError());
    ca::ctx().launches++;
    return CA_OK;
}

template <bool E, bool S>
static int launch_pipe(const RowBlockParams& p, int32_t nitems, cudaStream_t s) {
    auto kern = rowblock_pipe_kernel<E, S>;
    CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes));
    kern<<<(unsigned)nitems, THREADS, (size_t)p.smem_bytes, s>>>(p);
    CA_CUDA(cudaGetLastError());
    ca::ctx().launches++;
    return CA_OK;
}

// kind: 0 = resident items (16-bit counters), 1 = streaming items (16-bit), 2 = streaming items (32-bit),
//       3 = resident items through the pipelined kernel
int launch_rowblock(const RowBlockParams& p, int32_t nitems, int kind, cudaStream_t s) {
    if (nitems <= 0) return CA_OK;
    const bool e = p.ecc != nullptr, sm = p.sim != nullptr;
    if (kind == 0) {
        if (e && sm) return launch_one<false, true, true, true>(p, nitems, s);
        if (e) return launch_one<false, true, true, false>(p, nitems, s);
        return launch_one<false, true, false, true>(p, nitems, s);
    }
    if (kind == 3) {
        if (e && sm) return launch_pipe<true, true>(p, nitems, s);
        if (e) return launch_pipe<true, false>(p, nitems, s);
        return launch_pipe<false, true>(p, nitems, s);
    }
    if (kind == 1) {
        if (e && sm) return launch_one<false, false, true, true>(p, nitems, s);
        if (e) return launch_one<false, false, true, false>(p, nitems, s);
        return launch_one<false, false, false, true>(p, nitems, s);
    }
    if (e && sm) return launch_one<true, false, true, true>(p, nitems, s);
    if (e) return launch_one<true, false, true, false>(p, nitems, s);
    return launch_one<true, false, false, true>(p, nitems, s);
}

}  // namespace ca
