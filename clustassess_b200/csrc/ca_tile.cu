// ca_tile.cu -- the all-pairs ECS engine, second generation ("partner-tile" kernel, sm_100a).
//
// What it replaces: the U(U-1)/2 pair loop of weighted_element_consistency (R/ECS.R:1473-1490) with one
// disjointECS .Call per pair (src/calculate_ecs.cpp:22-46: contingency table, row/col sums, ratio table,
// per-element gather) and the pair loop of element_sim_matrix_flat_disjoint (R/ECS.R:953-959).
//
// Work item = (owner partition i, a set of COMPLETE clusters of i).  The cells of i are sorted by cluster
// (perm_i, ca_pack.cu), so a block that owns whole clusters owns whole ROWS of every contingency table T_ij.
// A thread owns a quad of consecutive cells of one cluster.  The block walks the partner partitions j in tiles
// of 8 (one 16-byte load per cell from the cell-major uint16 label matrix fetches 8 partners into registers)
// and, per PHASE (as many partners of the tile as fit in half of the shared memory):
//
//   fill    warp 0 initialises the phase's tables with TMA bulk copies completing on an mbarrier: row r of
//           partner j's table is a copy of the global array spack[j][b] = min(|B_j(b)|, 65535) << 16, so every
//           32-bit entry is born as (partner cluster size : 16 | count = 0 : 16);
//   hist    for each partner and cell: red.shared.add.u32 [entry], 1 -- with a CONSTANT 1 this is the SASS
//           instruction ATOMS.POPC.INC, which merges lanes that hit the same address in hardware (measured,
//           tools/microbench2.cu: 1.1 SM-cycles per warp instruction whether 32 lanes hit 32 banks or one
//           word; an add of a register value serialises instead).  Peaked label distributions -- the normal
//           case for similar partitions -- are therefore the FAST case and need no aggregation code;
//   gather  one shared load per (cell, partner) returns count and partner cluster size together;
//           ECS = T[a][b] / max(S_i[a], S_j[b])  (calculate_ecs.cpp:35) is accumulated in registers.
//
// Phases are double-buffered: while the block works on buffer p&1, the fill of buffer (p+1)&1 is in flight, and
// ONE __syncthreads per phase (between hist and gather) is the only block-wide barrier; the mbarrier of the
// next buffer also publishes the next phase's meta data.  Tables are row-major per partner with row stride
// K_j (rounded to 4), so a table takes R * K_j entries whatever the largest cluster count is.
// "Resident" items (<= 4096 positions) keep per-cell accumulators in registers over ALL partners and add to
// ecc once; clusters larger than that are "streaming" items: one row, both passes stream the cells from global
// memory, accumulators are flushed per phase; with more than 65535 cells the entries are plain 32-bit counts
// and the partner sizes come from global memory.
//
// The division is T * r with r = 1/max from rcp.approx.ftz.f64 + one Newton step (relative error < 1e-13);
// the batched path has to meet 1e-9 absolute, the per-pair entry points (ca_pair.cu) keep the IEEE division.
#include "ca_internal.h"

namespace ca {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int NT = TL_THREADS;
static constexpr int CAP = TL_CAP;
static_assert(TL_CPT == 4, "a thread owns one quad of cells");

// shared-memory map (byte offsets from the dynamic shared base)
static constexpr int SM_MBAR = 0;       // two mbarriers (one per buffer): tables initialised + meta published
static constexpr int SM_META = 16;      // two Meta records of 128 bytes
static constexpr int META_J0 = 0, META_G0 = 4, META_CNT = 8, META_TB = 16, META_RB = 48, META_KK = 80;
static constexpr int SM_SIMACC = 272;   // [2][8] double: sum of ECS over the block's cells, per buffer and partner
static constexpr int SM_DYN = 512;      // two table buffers
static constexpr uint32_t KK_BIG = 1u << 30;  // partner has a cluster with >= 65535 cells

// ---- shared-space accessors ---------------------------------------------------------------------------------
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
// the immediate 1 matters: ptxas emits ATOMS.POPC.INC (same-address lanes merged in hardware) only for it
__device__ __forceinline__ void red_inc(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }
__device__ __forceinline__ void red_add_f64(uint32_t a, double v) { asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// ---- TMA bulk copies completing on an mbarrier (SASS UBLKCP) ---------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();  // a lost copy must fail loudly, never hang the GPU
    }
}

__device__ __forceinline__ uint4 ldg_u4(const uint16_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// T / max(sa, sb): reciprocal seed + one Newton step, then one multiply
__device__ __forceinline__ double ecs_of(uint32_t count, uint32_t sa, uint32_t sb) {
    const double mx = (double)max(sa, sb);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(mx));
    r = fma(fma(-mx, r, 1.0), r, r);
    return (double)count * r;
}

template <int JT>
__device__ __forceinline__ uint32_t label_of(const uint4& v) {
    const uint32_t w = (JT >> 1) == 0 ? v.x : (JT >> 1) == 1 ? v.y : (JT >> 1) == 2 ? v.z : v.w;
    return (JT & 1) ? (w >> 16) : (w & 0xffffu);
}

// labels of the 8 partners [j0, j0+8) for a quad of cells: one 16-byte load per cell
__device__ __forceinline__ void load_quad(const uint16_t* __restrict__ lab16j, int mpad, int4 cell, uint4 (&L)[4]) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    L[0] = cell.x >= 0 ? ldg_u4(lab16j + (int64_t)cell.x * mpad) : z;
    L[1] = cell.y >= 0 ? ldg_u4(lab16j + (int64_t)cell.y * mpad) : z;
    L[2] = cell.z >= 0 ? ldg_u4(lab16j + (int64_t)cell.z * mpad) : z;
    L[3] = cell.w >= 0 ? ldg_u4(lab16j + (int64_t)cell.w * mpad) : z;
}

// ---- phase planning (warp 0) -----------------------------------------------------------------------------------
struct Planner {    // cursor of the next phase to plan; lives in warp 0's registers
    int j0, g0;     // tile and first tile slot
    int kreg;       // lanes 0..7: K_j | flags of tile j0's partners
    int knext;      // ... of tile j0 + 8
};

// Plans the phase at the cursor into buffer `buf`: how many partners of the tile fit (R rows of K_j entries each,
// K_j rounded to 4), their table addresses, and the bulk copies that initialise them.  Past the last partner it
// publishes cnt = 0.  Everything written here is published by the arrive on the buffer's mbarrier.
template <bool WIDE>
__device__ __forceinline__ void plan_phase(const TileParams& P, uint32_t sbase, uint32_t buf, uint32_t bufaddr, int half, int lane,
                                           Planner& pl, int jbeg, int jend, int R) {
    const uint32_t meta = sbase + SM_META + buf * 128u, bar = sbase + SM_MBAR + buf * 8u;
    if (pl.j0 >= jend) {
        if (lane == 0) {
            sts_u32(meta + META_CNT, 0u);
            mbar_arrive(bar);
        }
        return;
    }
    const int thi = min(8, jend - pl.j0);
    const int q = pl.g0 + lane;                                 // tile slot served by meta entry `lane`
    const uint32_t kk = (uint32_t)__shfl_sync(FULL, pl.kreg, q & 7);
    const bool ok = lane < 8 && q < thi;
    const uint32_t rowb = ok ? (((kk & 0xffffu) + 3u) & ~3u) * 4u : 0u;  // bytes of one table row
    const uint32_t tbytes = rowb * (uint32_t)R;
    uint32_t pre = tbytes;                                      // inclusive prefix of the bytes needed
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, pre, o);
        if (lane >= o) pre += v;
    }
    const unsigned fm = __ballot_sync(FULL, ok && pre <= (uint32_t)half);
    const int cnt = __popc(fm & 0xffu);                         // >= 1: the planner bounds R so that one partner always fits
    if (cnt == 0) __trap();                                     // cannot happen; never hang the GPU
    const uint32_t ttot = __shfl_sync(FULL, pre, cnt - 1);
    const uint32_t tb = bufaddr + pre - tbytes;
    if (lane < cnt) {
        sts_u32(meta + META_TB + lane * 4, tb);
        sts_u32(meta + META_RB + lane * 4, rowb);
        sts_u32(meta + META_KK + lane * 4, kk);
    }
    if (lane == 0) {
        sts_u32(meta + META_J0, (uint32_t)pl.j0);
        sts_u32(meta + META_G0, (uint32_t)pl.g0);
        sts_u32(meta + META_CNT, (uint32_t)cnt);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the block's generic accesses before the async writes
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, ttot);
        if (WIDE) bulk_g2s(bufaddr, P.zeros, ttot, bar);
    }
    __syncwarp();
    if (!WIDE) {
        for (int t = 0; t < cnt; t++) {  // partner t: R row copies, a lane each
            const uint32_t tbt = __shfl_sync(FULL, tb, t), rbt = __shfl_sync(FULL, rowb, t);
            const uint32_t* src = P.spack + (int64_t)(pl.j0 + pl.g0 + t) * P.kp;
            for (int r = lane; r < R; r += 32) bulk_g2s(tbt + (uint32_t)r * rbt, src, rbt, bar);
        }
    }
    // advance the cursor
    pl.g0 += cnt;
    if (pl.g0 >= thi) {
        pl.j0 += 8;
        pl.g0 = 0;
        pl.kreg = pl.knext;
        pl.knext = (lane < 8 && pl.j0 + 8 < jend) ? __ldg(P.kinfo + pl.j0 + 8 + lane) : 0;
    }
}

__device__ __forceinline__ void planner_init(const TileParams& P, Planner& pl, int jbeg, int jend, int lane) {
    pl.j0 = jbeg & ~7;
    pl.g0 = jbeg - pl.j0;
    pl.kreg = lane < 8 ? __ldg(P.kinfo + pl.j0 + lane) : 0;
    pl.knext = (lane < 8 && pl.j0 + 8 < jend) ? __ldg(P.kinfo + pl.j0 + 8 + lane) : 0;
}

struct Phase {
    int j0, g0, cnt;
};

// ---- per-slot histogram / gather for a quad (tile slot JT is a compile-time constant so that the label comes
//      straight out of its register; slots outside the phase are skipped by a block-uniform branch) -------------
template <int JT>
__device__ __forceinline__ void hist_slot(uint32_t meta, const Phase ph, const uint4 (&L)[4], uint32_t rowidx, int nvalid) {
    if (JT < ph.g0 || JT >= ph.g0 + ph.cnt) return;
    const int q = JT - ph.g0;
    const uint32_t rb = lds_u32(meta + META_TB + q * 4) + rowidx * lds_u32(meta + META_RB + q * 4);
#pragma unroll
    for (int c = 0; c < 4; c++)
        if (c < nvalid) red_inc(rb + label_of<JT>(L[c]) * 4u);
}
__device__ __forceinline__ void hist_quad(uint32_t meta, const Phase ph, const uint4 (&L)[4], uint32_t rowidx, int nvalid) {
    hist_slot<0>(meta, ph, L, rowidx, nvalid);
    hist_slot<1>(meta, ph, L, rowidx, nvalid);
    hist_slot<2>(meta, ph, L, rowidx, nvalid);
    hist_slot<3>(meta, ph, L, rowidx, nvalid);
    hist_slot<4>(meta, ph, L, rowidx, nvalid);
    hist_slot<5>(meta, ph, L, rowidx, nvalid);
    hist_slot<6>(meta, ph, L, rowidx, nvalid);
    hist_slot<7>(meta, ph, L, rowidx, nvalid);
}

template <int JT, bool WIDE, bool WANT_SIM>
__device__ __forceinline__ void gather_slot(const TileParams& P, uint32_t meta, uint32_t simacc, const Phase ph, const uint4 (&L)[4],
                                            uint32_t rowidx, uint32_t sa, int nvalid, int lane, double (&acc)[4]) {
    if (JT < ph.g0 || JT >= ph.g0 + ph.cnt) return;
    const int q = JT - ph.g0;
    const uint32_t rb = lds_u32(meta + META_TB + q * 4) + rowidx * lds_u32(meta + META_RB + q * 4);
    const uint32_t kk = lds_u32(meta + META_KK + q * 4);
    const double wj = __ldg(P.w + ph.j0 + JT);
    const int32_t* szg = P.sizes + (int64_t)(ph.j0 + JT) * P.kp;
    double es = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (c < nvalid) {
            const uint32_t lab = label_of<JT>(L[c]);
            const uint32_t e = lds_u32(rb + lab * 4u);
            uint32_t count, sb;
            if (WIDE) {
                count = e;
                sb = (uint32_t)__ldg(szg + lab);
            } else {
                count = e & 0xffffu;
                sb = e >> 16;
                if ((kk & KK_BIG) && sb == 0xffffu) sb = (uint32_t)__ldg(szg + lab);
            }
            const double v = ecs_of(count, sa, sb);
            acc[c] = fma(v, wj, acc[c]);
            if (WANT_SIM) es += v;
        }
    }
    if (WANT_SIM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(FULL, es, o);
        if (lane == 0 && es != 0.0) red_add_f64(simacc + q * 8u, es);
    }
}
template <bool WIDE, bool WANT_SIM>
__device__ __forceinline__ void gather_quad(const TileParams& P, uint32_t meta, uint32_t simacc, const Phase ph, const uint4 (&L)[4],
                                            uint32_t rowidx, uint32_t sa, int nvalid, int lane, double (&acc)[4]) {
    gather_slot<0, WIDE, WANT_SIM>(P, meta, simacc, ph, L, rowidx, sa, nvalid, lane, acc);
    gather_slot<1, WIDE, WANT_SIM>(P, meta, simacc, ph, L, rowidx, sa, nvalid, lane, acc);
    gather_slot<2, WIDE, WANT_SIM>(P, meta, simacc, ph, L, rowidx, sa, nvalid, lane, acc);
    gather_slot<3, WIDE, WANT_SIM>(P, meta, simacc, ph, L, rowidx, sa, nvalid, lane, acc);
    gather_slot<4, WIDE, WANT_SIM>(P, meta, simacc, ph, L, rowidx, sa, nvalid, lane, acc);
    gather_slot<5, WIDE, WANT_SIM>(P, meta, simacc, ph, L, rowidx, sa, nvalid, lane, acc);
    gather_slot<6, WIDE, WANT_SIM>(P, meta, simacc, ph, L, rowidx, sa, nvalid, lane, acc);
    gather_slot<7, WIDE, WANT_SIM>(P, meta, simacc, ph, L, rowidx, sa, nvalid, lane, acc);
}

// the per-partner ECS sums of a finished phase go to the ECS matrix (called by 8 threads after a block barrier)
__device__ __forceinline__ void flush_sim(const TileParams& P, uint32_t simacc, const Phase ph, int m, int t) {
    if (t < ph.cnt) {
        const double v = lds_f64(simacc + t * 8u);
        if (v != 0.0) atomicAdd(P.sim + (int64_t)m * P.m + (ph.j0 + ph.g0 + t), v);
        sts_f64(simacc + t * 8u, 0.0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One work item.  RESIDENT: <= CAP positions, R >= 1 complete clusters, accumulators in registers over all
// partners.  Otherwise ONE cluster with more than CAP cells, streamed in chunks of CAP positions in both passes.
// ---------------------------------------------------------------------------------------------------------------
template <bool RESIDENT, bool WIDE, bool WANT_ECC, bool WANT_SIM>
__device__ __forceinline__ void run_item(const TileParams& P, const Item it, unsigned char* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = it.m, R = RESIDENT ? it.nrows : 1;
    const int jbeg = it.jbeg, jend = it.jend;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const int half = ((P.smem_bytes - SM_DYN) / 2) & ~127;
    const int slot = P.slot_of[m];
    const int32_t* perm = P.perm + (int64_t)slot * P.perm_stride + it.pos0;
    const int mpad = P.mpad, npos = it.npos;
    const double wscale = P.w[m] * P.inv_norm;

    int4 cell = make_int4(-1, -1, -1, -1);
    int nvalid = 0;
    uint32_t rowidx = 0, sa = (uint32_t)it.sa;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (RESIDENT) {
        if (tid * 4 < npos) cell = __ldg(reinterpret_cast<const int4*>(perm + tid * 4));
        nvalid = (cell.x >= 0) + (cell.y >= 0) + (cell.z >= 0) + (cell.w >= 0);
        if (R > 1 && cell.x >= 0) {
            const int a = (int)__ldg(P.lab16 + (int64_t)cell.x * mpad + m);
            rowidx = (uint32_t)__ldg(P.crow + (int64_t)slot * P.kp + a);
            sa = (uint32_t)__ldg(P.sizes + (int64_t)m * P.kp + a);
        }
    }
    if (tid == 0) {
        mbar_init(sbase + SM_MBAR, 1);
        mbar_init(sbase + SM_MBAR + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 16) sts_f64(sbase + SM_SIMACC + tid * 8u, 0.0);
    __syncthreads();
    Planner pl;
    if (warp == 0) {
        planner_init(P, pl, jbeg, jend, lane);
        plan_phase<WIDE>(P, sbase, 0u, sbase + SM_DYN, half, lane, pl, jbeg, jend, R);
    }

    uint4 L[4];
    int cur_j0 = -1;
    Phase prev;
    prev.j0 = prev.g0 = prev.cnt = 0;
    uint32_t p = 0;
    for (;; p++) {
        const uint32_t b = p & 1u;
        const uint32_t meta = sbase + SM_META + b * 128u, simacc = sbase + SM_SIMACC + b * 64u;
        mbar_wait(sbase + SM_MBAR + b * 8u, (p >> 1) & 1u);  // buffer b initialised, meta of phase p published
        Phase ph;
        ph.cnt = (int)lds_u32(meta + META_CNT);
        if (ph.cnt == 0) break;
        ph.j0 = (int)lds_u32(meta + META_J0);
        ph.g0 = (int)lds_u32(meta + META_G0);
        // ---- histogram ----
        if (RESIDENT) {
            if (ph.j0 != cur_j0) {
                cur_j0 = ph.j0;
                load_quad(P.lab16 + ph.j0, mpad, cell, L);
                if (ph.j0 + 8 < jend) {  // ask L2 for the next tile's sectors now; its loads then hit
                    const uint16_t* ln = P.lab16 + ph.j0 + 8;
                    if (cell.x >= 0) prefetch_l2(ln + (int64_t)cell.x * mpad);
                    if (cell.y >= 0) prefetch_l2(ln + (int64_t)cell.y * mpad);
                    if (cell.z >= 0) prefetch_l2(ln + (int64_t)cell.z * mpad);
                    if (cell.w >= 0) prefetch_l2(ln + (int64_t)cell.w * mpad);
                }
            }
            hist_quad(meta, ph, L, rowidx, nvalid);
        } else {
            for (int cb = tid * 4; cb < npos; cb += CAP) {
                cell = __ldg(reinterpret_cast<const int4*>(perm + cb));
                nvalid = (cell.x >= 0) + (cell.y >= 0) + (cell.z >= 0) + (cell.w >= 0);
                load_quad(P.lab16 + ph.j0, mpad, cell, L);
                hist_quad(meta, ph, L, 0u, nvalid);
            }
        }
        __syncthreads();  // the tables of phase p are complete; every thread has left phase p-1
        if (warp == 0) plan_phase<WIDE>(P, sbase, b ^ 1u, sbase + SM_DYN + (b ^ 1u) * (uint32_t)half, half, lane, pl, jbeg, jend, R);
        if (WANT_SIM && warp == 1) flush_sim(P, sbase + SM_SIMACC + (b ^ 1u) * 64u, prev, m, lane);
        // ---- gather ----
        if (RESIDENT) {
            gather_quad<WIDE, WANT_SIM>(P, meta, simacc, ph, L, rowidx, sa, nvalid, lane, acc);
        } else {
            for (int cb0 = 0; cb0 < npos; cb0 += CAP) {  // block-uniform trip count: gather_quad shuffles when WANT_SIM
                const int cb = cb0 + tid * 4;
                cell = make_int4(-1, -1, -1, -1);
                if (cb < npos) cell = __ldg(reinterpret_cast<const int4*>(perm + cb));
                nvalid = (cell.x >= 0) + (cell.y >= 0) + (cell.z >= 0) + (cell.w >= 0);
                load_quad(P.lab16 + ph.j0, mpad, cell, L);
                double a4[4] = {0.0, 0.0, 0.0, 0.0};
                gather_quad<WIDE, WANT_SIM>(P, meta, simacc, ph, L, 0u, sa, nvalid, lane, a4);
                if (WANT_ECC) {
                    if (cell.x >= 0) atomicAdd(P.ecc + cell.x, a4[0] * wscale);
                    if (cell.y >= 0) atomicAdd(P.ecc + cell.y, a4[1] * wscale);
                    if (cell.z >= 0) atomicAdd(P.ecc + cell.z, a4[2] * wscale);
                    if (cell.w >= 0) atomicAdd(P.ecc + cell.w, a4[3] * wscale);
                }
            }
        }
        prev = ph;
    }
    if (WANT_SIM) {
        __syncthreads();
        if (warp == 1) flush_sim(P, sbase + SM_SIMACC + ((p - 1u) & 1u) * 64u, prev, m, lane);
    }
    if (RESIDENT && WANT_ECC) {
        if (cell.x >= 0) atomicAdd(P.ecc + cell.x, acc[0] * wscale);
        if (cell.y >= 0) atomicAdd(P.ecc + cell.y, acc[1] * wscale);
        if (cell.z >= 0) atomicAdd(P.ecc + cell.z, acc[2] * wscale);
        if (cell.w >= 0) atomicAdd(P.ecc + cell.w, acc[3] * wscale);
    }
}

template <bool WANT_ECC, bool WANT_SIM>
__global__ void __launch_bounds__(NT, 1) tile_kernel(const TileParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Item it = P.items[blockIdx.x];
    if (it.jbeg >= it.jend) return;
    if (it.npos <= CAP)
        run_item<true, false, WANT_ECC, WANT_SIM>(P, it, smem);
    else if (it.pad_)
        run_item<false, true, WANT_ECC, WANT_SIM>(P, it, smem);
    else
        run_item<false, false, WANT_ECC, WANT_SIM>(P, it, smem);
}

// Shared memory one block asks for and the number of table rows a resident block may own so that at least one
// partner with kmax clusters fits into one of the two buffers (R rows of kmax 32-bit entries, kmax rounded to 4).
int tile_smem_budget(int32_t kmax, int32_t* max_rows_out) {
    Ctx& c = ctx();
    const size_t optin = c.smem_optin ? c.smem_optin : 227 * 1024;
    const int budget = (int)optin;
    const long long half = (((long long)budget - SM_DYN) / 2) & ~127LL;
    const long long rowb = (((long long)(kmax > 0 ? kmax : 1) + 3) & ~3LL) * 4;
    long long rows = half / rowb;
    if (rows > TL_MAX_ROWS) rows = TL_MAX_ROWS;
    if (max_rows_out) *max_rows_out = (int32_t)(rows < 0 ? 0 : rows);
    return budget;
}

template <bool E, bool S>
static int launch_tile_t(const TileParams& p, int32_t nitems, cudaStream_t s) {
    auto kern = tile_kernel<E, S>;
    CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes));
    kern<<<(unsigned)nitems, NT, (size_t)p.smem_bytes, s>>>(p);
    CA_CUDA(cudaGetLastError());
    ca::ctx().launches++;
    return CA_OK;
}

int launch_tile(const TileParams& p, int32_t nitems, cudaStream_t s) {
    if (nitems <= 0) return CA_OK;
    const bool e = p.ecc != nullptr, sm = p.sim != nullptr;
    if (e && sm) return launch_tile_t<true, true>(p, nitems, s);
    if (e) return launch_tile_t<true, false>(p, nitems, s);
    if (sm) return launch_tile_t<false, true>(p, nitems, s);
    return CA_OK;
}

}  // namespace ca
