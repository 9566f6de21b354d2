// ca_tile.cu -- the all-pairs ECS engine, second generation ("partner-tile" kernel, sm_100a).
//
// What it replaces: the U(U-1)/2 pair loop of weighted_element_consistency (R/ECS.R:1473-1490) with one
// disjointECS .Call per pair (src/calculate_ecs.cpp:22-46: contingency table, row/col sums, ratio table,
// per-element gather) and the pair loop of element_sim_matrix_flat_disjoint (R/ECS.R:953-959).
//
// Work item = (owner partition i, a set of COMPLETE clusters of i).  The cells of i are sorted by cluster
// (perm_i, ca_pack.cu), so a block that owns whole clusters owns whole ROWS of every contingency table T_ij.
// The block walks the partner partitions j in tiles of 8 (one 16-byte load per cell from the cell-major uint16
// label matrix fetches 8 partners) and, per phase (as many partners of the tile as fit in shared memory):
//
//   fill    one thread zero-fills the phase's tables and stages the partners' cluster sizes (as uint16, 0xFFFF
//           = "look it up in global memory") with TMA bulk copies completing on an mbarrier;
//   hist    lanes are mapped to PARTNERS, not to cells: lane = (cell group, partner).  Every partner has its
//           own table, so the lanes of one atomic instruction can collide on an address only within the
//           32/partners cell groups (<= 4-way for 8 partners) however peaked the label distribution is --
//           no run detection, no warp aggregation, a handful of instructions per 32 (pair, cell) units;
//   gather  lanes are mapped to cells (4 consecutive cells of one cluster per thread): count and partner
//           cluster size come from shared memory (loads broadcast, so collisions are free),
//           ECS = T[a][b] / max(S_i[a], S_j[b])  (calculate_ecs.cpp:35) is accumulated in registers.
//
// Per-partner tables are COLUMN-major, counter (label b, row r) at b*R + r (R = rows of this block), so a table
// takes K_j*R counters whatever the largest cluster count is, and the key is computed from the raw label.
// Counters are 16 bit (two per word; a resident block holds <= 4096 cells) unless a streaming item has more
// than 65535 cells.  "Resident" items (<= 4096 positions) keep per-cell accumulators in registers over ALL
// partners and add to ecc once; clusters larger than that are "streaming" items: one row, both passes stream
// the cells from global memory, accumulators are flushed per tile.
//
// The division is T * r with r = 1/max from rcp.approx.ftz.f64 + one Newton step (relative error < 1e-13);
// the batched path has to meet 1e-9 absolute, the per-pair entry points (ca_pair.cu) keep the IEEE division.
#include "ca_internal.h"

namespace ca {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int NT = TL_THREADS;
static constexpr int NWARP = NT / 32;
static constexpr int CAP = TL_CAP;
static_assert(TL_CPT == 4, "a thread owns one quad of cells");

// shared-memory map (byte offsets from the dynamic shared base)
static constexpr int SM_MBAR = 0;                  // one mbarrier: tables zeroed + sizes staged
static constexpr int SM_CNT = 8;                   // partners in this phase
static constexpr int SM_TB = 16;                   // [8] u32 shared address of partner q's table
static constexpr int SM_SZ = 48;                   // [8] u32 shared address of partner q's uint16 sizes
static constexpr int SM_KK = 80;                   // [8] u32 K_j | flags of partner q
static constexpr int SM_SIMACC = 128;              // [8] double: sum of ECS over the block's cells, per partner
static constexpr int SM_ROWQ = 192;                // [NT] u16 (+ 4 bytes of padding per 32): table row of every thread's quad
static constexpr int SM_TILE = SM_ROWQ + NT * 2 + (NT / 32) * 4;  // [8][CAP + 4] u16: labels of the current 8 partners, partner-major
static constexpr int TILE_ROW = CAP * 2 + 8;       // row stride: 2 words of padding shift consecutive partners by 2 banks
static constexpr int SM_DYN = SM_TILE + 8 * TILE_ROW;  // tables, then sizes
static constexpr uint32_t KK_BIG = 1u << 30;       // partner has a cluster with >= 65535 cells

// ---- shared-space accessors ---------------------------------------------------------------------------------
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
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void red_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add_f64(uint32_t a, double v) { asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// ---- TMA bulk copies completing on an mbarrier (SASS UBLKCP) ---------------------------------------------------
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

__device__ __forceinline__ void sts_v2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void lds_v2(uint32_t a, uint32_t& x, uint32_t& y) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
}

// Labels of 8 partners for a quad of cells (one 16-byte load per cell; padding cells read as 0xFFFF, which is
// never a label because K <= 65535), transposed in registers and parked partner-major in the thread's own 8-byte
// slot of every tile row: row q holds { cell0 | cell1 << 16, cell2 | cell3 << 16 } of partner j0 + q.
__device__ __forceinline__ void stage_quad(const uint16_t* __restrict__ lab16j, int mpad, int4 cell, uint32_t s_slot) {
    const uint4 padv = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    const uint4 a = cell.x >= 0 ? ldg_u4(lab16j + (int64_t)cell.x * mpad) : padv;
    const uint4 b = cell.y >= 0 ? ldg_u4(lab16j + (int64_t)cell.y * mpad) : padv;
    const uint4 c = cell.z >= 0 ? ldg_u4(lab16j + (int64_t)cell.z * mpad) : padv;
    const uint4 d = cell.w >= 0 ? ldg_u4(lab16j + (int64_t)cell.w * mpad) : padv;
    sts_v2(s_slot + 0 * TILE_ROW, __byte_perm(a.x, b.x, 0x5410), __byte_perm(c.x, d.x, 0x5410));
    sts_v2(s_slot + 1 * TILE_ROW, __byte_perm(a.x, b.x, 0x7632), __byte_perm(c.x, d.x, 0x7632));
    sts_v2(s_slot + 2 * TILE_ROW, __byte_perm(a.y, b.y, 0x5410), __byte_perm(c.y, d.y, 0x5410));
    sts_v2(s_slot + 3 * TILE_ROW, __byte_perm(a.y, b.y, 0x7632), __byte_perm(c.y, d.y, 0x7632));
    sts_v2(s_slot + 4 * TILE_ROW, __byte_perm(a.z, b.z, 0x5410), __byte_perm(c.z, d.z, 0x5410));
    sts_v2(s_slot + 5 * TILE_ROW, __byte_perm(a.z, b.z, 0x7632), __byte_perm(c.z, d.z, 0x7632));
    sts_v2(s_slot + 6 * TILE_ROW, __byte_perm(a.w, b.w, 0x5410), __byte_perm(c.w, d.w, 0x5410));
    sts_v2(s_slot + 7 * TILE_ROW, __byte_perm(a.w, b.w, 0x7632), __byte_perm(c.w, d.w, 0x7632));
}

__device__ __forceinline__ uint32_t rowq_off(uint32_t quad) { return SM_ROWQ + quad * 2u + (quad >> 5) * 4u; }

struct Phase {      // what every thread needs to know about the current phase
    int g0, cnt;    // partners [g0, g0+cnt) of the current 8-partner tile
};

// Warp 0 plans a phase: how many partners of the tile, starting at tile slot g0, fit into `avail` bytes of shared
// memory (tables first, contiguous, then the uint16 size rows), writes their addresses to the meta area and
// starts the bulk copies.  kreg (lanes 0..7) = K_j | flags of the tile's 8 partners.  ctr_bytes = 2 or 4.
__device__ __forceinline__ void plan_and_fill(const TileParams& P, uint32_t sbase, uint32_t dyn, int avail, int lane, int j0,
                                              int g0, int thi, int kreg, int R, int ctr_bytes) {
    const int q = g0 + lane;                                   // tile slot served by meta entry `lane`
    const uint32_t kk = (uint32_t)__shfl_sync(FULL, kreg, q & 7);
    const bool ok = lane < 8 && q < thi;
    const uint32_t K = ok ? (kk & 0xffffu) : 0u;
    const uint32_t tbytes = (K * (uint32_t)R * (uint32_t)ctr_bytes + 15u) & ~15u;
    const uint32_t sbytes = (K * 2u + 15u) & ~15u;
    uint32_t pre = tbytes + sbytes;                            // inclusive prefix of the bytes needed
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, pre, o);
        if (lane >= o) pre += v;
    }
    const unsigned fm = __ballot_sync(FULL, ok && pre <= (uint32_t)avail);
    const int cnt = __popc(fm & 0xffu);                        // >= 1: the planner bounds R so that one partner always fits
    if (cnt == 0) __trap();                                    // cannot happen; never hang the GPU
    const bool mine = lane < cnt;
    uint32_t tpre = mine ? tbytes : 0u, spre = mine ? sbytes : 0u;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, tpre, o), u = __shfl_up_sync(FULL, spre, o);
        if (lane >= o) {
            tpre += v;
            spre += u;
        }
    }
    const uint32_t ttot = __shfl_sync(FULL, tpre, 7), stot = __shfl_sync(FULL, spre, 7);
    const uint32_t bar = sbase + SM_MBAR;
    if (mine) {
        sts_u32(sbase + SM_TB + lane * 4, dyn + tpre - tbytes);
        sts_u32(sbase + SM_SZ + lane * 4, dyn + ttot + spre - sbytes);
        sts_u32(sbase + SM_KK + lane * 4, kk);
    }
    if (lane == 0) {
        sts_u32(sbase + SM_CNT, (uint32_t)cnt);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the block's generic accesses before the async writes
        mbar_expect_tx(bar, ttot + stot);
        bulk_g2s(dyn, P.zeros, ttot, bar);
    }
    __syncwarp();
    if (mine) bulk_g2s(dyn + ttot + spre - sbytes, P.sizes16 + (int64_t)(j0 + q) * P.kp, sbytes, bar);
}

// histogram lane geometry for a phase of cnt partners: NG = 2^lg partner slots, 32/NG cell groups
struct HistGeom {
    int lg, jj, cg;
    bool on;
    uint32_t tb;
};
__device__ __forceinline__ HistGeom hist_geom(uint32_t sbase, int cnt, int lane) {
    HistGeom h;
    h.lg = cnt > 4 ? 3 : cnt > 2 ? 2 : cnt > 1 ? 1 : 0;
    h.jj = lane & ((1 << h.lg) - 1);
    h.cg = lane >> h.lg;
    h.on = h.jj < cnt;
    h.tb = lds_u32(sbase + SM_TB + (h.on ? h.jj : 0) * 4);
    return h;
}
template <bool WIDE>
__device__ __forceinline__ void tbl_inc(uint32_t tb, uint32_t key) {
    if (WIDE)
        red_add(tb + key * 4u, 1u);
    else  // two 16-bit counters per word; a block never counts past 65535, so the halves cannot carry
        red_add(tb + ((key * 2u) & ~3u), (key & 1u) * 0xffffu + 1u);
}
template <bool WIDE>
__device__ __forceinline__ uint32_t tbl_get(uint32_t tb, uint32_t key) {
    return WIDE ? lds_u32(tb + key * 4u) : lds_u16(tb + key * 2u);
}

// gather of the phase's partners for this thread's quad; the labels come back from the thread's own tile slots
template <bool WIDE, bool WANT_SIM>
__device__ __forceinline__ void gather_quad(const TileParams& P, uint32_t sbase, const Phase ph, int j0, uint32_t s_slot, uint32_t R,
                                            uint32_t rowidx, uint32_t sa, int nvalid, int lane, double (&acc)[4]) {
#pragma unroll 2
    for (int q = 0; q < ph.cnt; q++) {
        const int jt = ph.g0 + q;
        const uint32_t tb = lds_u32(sbase + SM_TB + q * 4), sz = lds_u32(sbase + SM_SZ + q * 4), kk = lds_u32(sbase + SM_KK + q * 4);
        const double wj = __ldg(P.w + j0 + jt);
        uint32_t p01, p23;
        lds_v2(s_slot + (uint32_t)jt * TILE_ROW, p01, p23);
        const uint32_t lab[4] = {p01 & 0xffffu, p01 >> 16, p23 & 0xffffu, p23 >> 16};
        double es = 0.0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c < nvalid) {
                const uint32_t count = tbl_get<WIDE>(tb, lab[c] * R + rowidx);
                uint32_t sb = lds_u16(sz + lab[c] * 2u);
                if ((kk & KK_BIG) && sb == 0xffffu) sb = (uint32_t)__ldg(P.sizes + (int64_t)(j0 + jt) * P.kp + lab[c]);
                const double e = ecs_of(count, sa, sb);
                acc[c] = fma(e, wj, acc[c]);
                if (WANT_SIM) es += e;
            }
        }
        if (WANT_SIM) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(FULL, es, o);
            if (lane == 0 && es != 0.0) red_add_f64(sbase + SM_SIMACC + q * 8u, es);
        }
    }
}

// after a phase's gather: add the per-partner ECS sums to the ECS matrix and clear them
__device__ __forceinline__ void flush_sim(const TileParams& P, uint32_t sbase, const Phase ph, int m, int j0, int tid) {
    __syncthreads();
    if (tid < ph.cnt) {
        const double v = lds_f64(sbase + SM_SIMACC + tid * 8u);
        if (v != 0.0) atomicAdd(P.sim + (int64_t)m * P.m + (j0 + ph.g0 + tid), v);
        sts_f64(sbase + SM_SIMACC + tid * 8u, 0.0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// resident item: <= CAP positions, R >= 1 complete clusters; per-cell accumulators live in registers
// ---------------------------------------------------------------------------------------------------------------
template <bool WANT_ECC, bool WANT_SIM>
__device__ __forceinline__ void resident_item(const TileParams& P, const Item it, unsigned char* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = it.m, R = it.nrows;
    const int jbeg = it.jbeg, jend = it.jend;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t s_tile = sbase + SM_TILE, s_slot = s_tile + tid * 8u, dyn = sbase + SM_DYN;
    const int avail = P.smem_bytes - SM_DYN;
    const int slot = P.slot_of[m];
    const int32_t* perm = P.perm + (int64_t)slot * P.perm_stride + it.pos0;
    const int mpad = P.mpad;

    int4 cell = make_int4(-1, -1, -1, -1);
    if (tid * 4 < it.npos) cell = __ldg(reinterpret_cast<const int4*>(perm + tid * 4));
    const int nvalid = (cell.x >= 0) + (cell.y >= 0) + (cell.z >= 0) + (cell.w >= 0);
    uint32_t rowidx = 0, sa = (uint32_t)it.sa;
    if (R > 1 && cell.x >= 0) {
        const int a = (int)__ldg(P.lab16 + (int64_t)cell.x * mpad + m);
        rowidx = (uint32_t)__ldg(P.crow + (int64_t)slot * P.kp + a);
        sa = (uint32_t)__ldg(P.sizes + (int64_t)m * P.kp + a);
    }
    sts_u16(sbase + rowq_off((uint32_t)tid), rowidx);
    if (tid == 0) {
        mbar_init(sbase + SM_MBAR, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 8) sts_f64(sbase + SM_SIMACC + tid * 8u, 0.0);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};

    const int jfirst = jbeg & ~7;
    int kreg = (warp == 0 && lane < 8) ? __ldg(P.kinfo + jfirst + lane) : 0;
    uint32_t par = 0;
    for (int j0 = jfirst; j0 < jend; j0 += 8) {
        const int tlo = max(0, jbeg - j0), thi = min(8, jend - j0);
        __syncthreads();  // every thread has finished with the previous tile and its tables
        stage_quad(P.lab16 + j0, mpad, cell, s_slot);
        int knext = 0;
        if (j0 + 8 < jend) {  // next tile: ask L2 for the sectors now, the loads one tile later then hit
            const uint16_t* ln = P.lab16 + j0 + 8;
            if (cell.x >= 0) prefetch_l2(ln + (int64_t)cell.x * mpad);
            if (cell.y >= 0) prefetch_l2(ln + (int64_t)cell.y * mpad);
            if (cell.z >= 0) prefetch_l2(ln + (int64_t)cell.z * mpad);
            if (cell.w >= 0) prefetch_l2(ln + (int64_t)cell.w * mpad);
            if (warp == 0 && lane < 8) knext = __ldg(P.kinfo + j0 + 8 + lane);
        }
        for (int g0 = tlo; g0 < thi;) {
            if (g0 != tlo) __syncthreads();  // previous phase's tables are free
            if (warp == 0) plan_and_fill(P, sbase, dyn, avail, lane, j0, g0, thi, kreg, R, 2);
            __syncthreads();  // tile and phase meta visible
            Phase ph;
            ph.g0 = g0;
            ph.cnt = (int)lds_u32(sbase + SM_CNT);
            const HistGeom hg = hist_geom(sbase, ph.cnt, lane);
            mbar_wait(sbase + SM_MBAR, par);
            par ^= 1u;
            // ---- histogram: lane = (cell group cg, partner slot jj); the groups start at staggered offsets inside
            //      their ranges so that the 32 lanes of a step read 32 different banks of the tile ----
            {
                static_assert(CAP == 4096, "lgc assumes 4096 positions");
                const int lgc = 7 + hg.lg;  // log2(cells per group) = log2(CAP / (32 >> lg))
                const uint32_t cmask = (1u << lgc) - 1u;
                const uint32_t tl = s_tile + (uint32_t)(g0 + hg.jj) * TILE_ROW;
                const uint32_t cbase = (uint32_t)hg.cg << lgc, stag = 2u * ((uint32_t)(hg.cg & 1) + ((uint32_t)(hg.cg >> 1) << (hg.lg + 1)));
                if (R > 1) {
#pragma unroll 4
                    for (uint32_t t = warp; t <= cmask; t += NWARP) {
                        const uint32_t cp = cbase + ((t + stag) & cmask);
                        const uint32_t lab = lds_u16(tl + cp * 2u);
                        const uint32_t ro = lds_u16(sbase + rowq_off(cp >> 2));
                        if (hg.on && lab != 0xffffu) tbl_inc<false>(hg.tb, lab * (uint32_t)R + ro);
                    }
                } else {
#pragma unroll 4
                    for (uint32_t t = warp; t <= cmask; t += NWARP) {
                        const uint32_t cp = cbase + ((t + stag) & cmask);
                        const uint32_t lab = lds_u16(tl + cp * 2u);
                        if (hg.on && lab != 0xffffu) tbl_inc<false>(hg.tb, lab);
                    }
                }
            }
            __syncthreads();
            // ---- gather: lane = quad of cells ----
            gather_quad<false, WANT_SIM>(P, sbase, ph, j0, s_slot, (uint32_t)R, rowidx, sa, nvalid, lane, acc);
            if (WANT_SIM) flush_sim(P, sbase, ph, m, j0, tid);
            g0 += ph.cnt;
        }
        kreg = knext;
    }
    if (WANT_ECC) {
        const double wscale = P.w[m] * P.inv_norm;
        if (cell.x >= 0) atomicAdd(P.ecc + cell.x, acc[0] * wscale);
        if (cell.y >= 0) atomicAdd(P.ecc + cell.y, acc[1] * wscale);
        if (cell.z >= 0) atomicAdd(P.ecc + cell.z, acc[2] * wscale);
        if (cell.w >= 0) atomicAdd(P.ecc + cell.w, acc[3] * wscale);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// streaming item: ONE cluster with more than CAP cells (one table row); both passes stream the cells from
// global memory, per-cell sums are added to ecc once per phase
// ---------------------------------------------------------------------------------------------------------------
template <bool WIDE, bool WANT_ECC, bool WANT_SIM>
__device__ __forceinline__ void streaming_item(const TileParams& P, const Item it, unsigned char* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = it.m;
    const int jbeg = it.jbeg, jend = it.jend;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t s_slot = sbase + SM_TILE + tid * 8u, dyn = sbase + SM_DYN;
    const int avail = P.smem_bytes - SM_DYN;
    const int slot = P.slot_of[m];
    const int32_t* perm = P.perm + (int64_t)slot * P.perm_stride + it.pos0;
    const int mpad = P.mpad, npos = it.npos;
    const uint32_t sa = (uint32_t)it.sa;
    const double wscale = P.w[m] * P.inv_norm;
    if (tid == 0) {
        mbar_init(sbase + SM_MBAR, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 8) sts_f64(sbase + SM_SIMACC + tid * 8u, 0.0);
    const int jfirst = jbeg & ~7;
    int kreg = (warp == 0 && lane < 8) ? __ldg(P.kinfo + jfirst + lane) : 0;
    uint32_t par = 0;
    for (int j0 = jfirst; j0 < jend; j0 += 8) {
        const int tlo = max(0, jbeg - j0), thi = min(8, jend - j0);
        int knext = 0;
        if (warp == 0 && lane < 8 && j0 + 8 < jend) knext = __ldg(P.kinfo + j0 + 8 + lane);
        for (int g0 = tlo; g0 < thi;) {
            __syncthreads();  // previous phase's tables are free
            if (warp == 0) plan_and_fill(P, sbase, dyn, avail, lane, j0, g0, thi, kreg, 1, WIDE ? 4 : 2);
            __syncthreads();
            Phase ph;
            ph.g0 = g0;
            ph.cnt = (int)lds_u32(sbase + SM_CNT);
            const HistGeom hg = hist_geom(sbase, ph.cnt, lane);
            mbar_wait(sbase + SM_MBAR, par);
            par ^= 1u;
            // ---- histogram: a warp step covers 32 >> lg consecutive positions x 2^lg partners ----
            {
                const int ncg = 32 >> hg.lg;
                const uint16_t* lcol = P.lab16 + j0 + g0 + hg.jj;
                constexpr int U = 4;
                for (int t0 = warp * U; t0 * ncg < npos; t0 += NWARP * U) {
                    int cid[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int p = (t0 + u) * ncg + hg.cg;
                        cid[u] = p < npos ? __ldg(perm + p) : -1;
                    }
                    uint32_t lab[U];
#pragma unroll
                    for (int u = 0; u < U; u++) lab[u] = (hg.on && cid[u] >= 0) ? (uint32_t)__ldg(lcol + (int64_t)cid[u] * mpad) : 0xffffu;
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (lab[u] != 0xffffu) tbl_inc<WIDE>(hg.tb, lab[u]);
                }
            }
            __syncthreads();
            // ---- gather: chunks of CAP positions, a quad per thread (labels parked in the thread's own tile slots) ----
            for (int cb = 0; cb < npos; cb += CAP) {
                const int p = cb + tid * 4;
                int4 cell = make_int4(-1, -1, -1, -1);
                if (p < npos) cell = __ldg(reinterpret_cast<const int4*>(perm + p));
                const int nvalid = (cell.x >= 0) + (cell.y >= 0) + (cell.z >= 0) + (cell.w >= 0);
                if (nvalid > 0) stage_quad(P.lab16 + j0, mpad, cell, s_slot);
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                gather_quad<WIDE, WANT_SIM>(P, sbase, ph, j0, s_slot, 1u, 0u, sa, nvalid, lane, acc);
                if (WANT_ECC) {
                    if (cell.x >= 0) atomicAdd(P.ecc + cell.x, acc[0] * wscale);
                    if (cell.y >= 0) atomicAdd(P.ecc + cell.y, acc[1] * wscale);
                    if (cell.z >= 0) atomicAdd(P.ecc + cell.z, acc[2] * wscale);
                    if (cell.w >= 0) atomicAdd(P.ecc + cell.w, acc[3] * wscale);
                }
            }
            if (WANT_SIM) flush_sim(P, sbase, ph, m, j0, tid);
            g0 += ph.cnt;
        }
        kreg = knext;
    }
}

template <bool WANT_ECC, bool WANT_SIM>
__global__ void __launch_bounds__(NT, 1) tile_kernel(const TileParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Item it = P.items[blockIdx.x];
    if (it.jbeg >= it.jend) return;
    if (it.npos <= CAP)
        resident_item<WANT_ECC, WANT_SIM>(P, it, smem);
    else if (it.pad_)
        streaming_item<true, WANT_ECC, WANT_SIM>(P, it, smem);
    else
        streaming_item<false, WANT_ECC, WANT_SIM>(P, it, smem);
}

// Shared memory one block asks for and the number of table rows a resident block may own so that at least one
// partner with kmax clusters fits (tables K*R 16-bit counters + K 16-bit sizes, both rounded to 16 bytes).
int tile_smem_budget(int32_t kmax, int32_t* max_rows_out) {
    Ctx& c = ctx();
    const size_t optin = c.smem_optin ? c.smem_optin : 227 * 1024;
    const int budget = (int)optin;
    const long long avail = (long long)budget - SM_DYN - 32;
    long long rows = avail / (2LL * (kmax > 0 ? kmax : 1)) - 1;
    if (rows > TL_MAX_ROWS) rows = TL_MAX_ROWS;
    // a streaming item needs one row of 32-bit counters: 6 bytes per cluster
    if (6LL * kmax + 32 > (long long)budget - SM_TILE) rows = 0;
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
