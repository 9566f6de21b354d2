// ca_tile.cu -- the all-pairs ECS engine: the persistent "wave" kernel (sm_100a).
//
// What it replaces: the U(U-1)/2 pair loop of weighted_element_consistency (R/ECS.R:1473-1490) with one
// disjointECS .Call per pair (src/calculate_ecs.cpp:22-46: contingency table, row/col sums, ratio table,
// per-element gather), the pair loop of element_sim_matrix_flat_disjoint (R/ECS.R:953-959) and the loop of
// element_agreement's flat branch (R/ECS.R:1639-1647).
//
// ITEM = a set of COMPLETE clusters of an owner partition i (host planner, ca_plan.cpp).  The cells of i are
// sorted by cluster (perm_i, ca_pack.cu), so a block that works on an item owns whole ROWS of every contingency
// table T_ij.
//
// WORK UNIT = (item, span of 32 partner partitions = two 16-partner super-tiles).  Work units are numbered span-major
// and the persistent blocks (one per SM) take them in that order, first come first served, so all SMs sweep the partner axis together and the
// labels of the current super-tiles -- dense 32 (N+1)-byte slabs of the super-tile-major uint16 label array,
// lab16_index() -- are fetched from HBM once and then served by L2 to every owner partition.
//
// PHASE = as many partners of the current 8-partner tile as fit into one of the two shared-memory buffers:
//   fill    the producer warp initialises the phase's tables with TMA bulk copies completing on the buffer's
//           `full` mbarrier: partner j's R-row table is ONE copy of the first R rows of j's table template in global
//           memory (row b-th entry = min(|B_j(b)|, 65535) << 16), so every 32-bit entry is born as
//           (partner cluster size : 16 | count = 0 : 16); the same mbarrier publishes the phase's meta record;
//   hist    for each partner and cell: red.shared.add.u32 [entry], 1 -- with a CONSTANT 1 this is the SASS
//           instruction ATOMS.POPC.INC, which merges lanes that hit the same address in hardware (measured,
//           tools/microbench2.cu: 1.1 SM-cycles per warp instruction whether 32 lanes hit 32 banks or one
//           word; an add of a register value serialises instead);
//   sync    ONE named barrier among the consumer warps;
//   gather  one shared load per (cell, partner) returns count and partner cluster size together;
//           ECS = T[a][b] / max(S_i[a], S_j[b])  (calculate_ecs.cpp:35) is accumulated in registers; the
//           consumer warps then release the buffer through its `empty` mbarrier.
// Tables are row-major per partner with row stride K_j + 1 (rounded to 4).  Column K_j is the PAD COLUMN: padding
// positions of the cluster-sorted order read the pad cell N, whose label in partition j is K_j, so the inner loops
// carry no validity tests.  The bodies of histogram and gather are instantiated per tile slot (label word and half are
// compile-time, the slot's meta entries sit at immediate offsets) and entered Duff-style at the phase's first slot.
//
// Two kinds of items, two launches (own block size and register allocation each):
//   RESIDENT (<= TL_CAP positions, R >= 1 complete clusters): a consumer thread owns a quad of consecutive positions
//     of one cluster.  It fetches the 16 labels of a cell for a whole super-tile with ONE 32-byte load (LDG.E.256) and
//     -- CA_TL_PREFETCH -- does so one tile AHEAD into a second register set: while the upper tile of a super-tile is
//     being processed the next super-tile's sectors are already in flight, and during a work unit's last tile so are
//     the next work unit's cell ids, quad info and first sectors (the producer publishes the next unit's coordinates
//     in the work-unit record).  The per-cell sums go to ecc once per work unit.
//   STREAMING (one cluster with more than TL_CAP cells, one table row): TL_STR_THREADS threads, cells in chunks of a quad per
//     thread; items of at most two chunks keep both chunks' labels in registers between histogram and gather, longer ones
//     stream the cells from global memory in both passes; labels are loaded per 8-partner tile (16 bytes per cell), sums
//     are added per phase.  More than 65535 cells: plain 32-bit counters, partner sizes from global memory.
//
// The ECS matrix (mean of a pair's ECS) has two routes, chosen per job by the host (TileParams::sim_scan): per (cell, partner) in
// the gather, or -- partitions with few clusters -- summed over the non-empty table entries after the histogram (scan_sim), in
// which case element_sim_matrix runs no gather at all.  A third set of instantiations (RT, TileParams::rt) serves jobs with few
// but huge clusters: one division per table entry into a table of doubles, and a gather that only adds (convert_tables).
//
// The division is T * r with r = 1/max from rcp.approx.ftz.f64 + one Newton step (relative error < 1e-13);
// the batched path has to meet 1e-9 absolute, the per-pair entry points (ca_pair.cu) keep the IEEE division.
// Sums leave the kernel as 64-bit fixed-point integers (ca_internal.h): integer atomics make the result independent
// of the order in which blocks, warps and GPUs add -- two runs give identical bits.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "ca_internal.h"

namespace ca {

#ifndef CA_BACKOFF_NS
#define CA_BACKOFF_NS 200  // the producer's sleep between probes of a buffer that is still in use (tuning builds)
#endif
// Checked builds (tools/build_variant.sh checked -DCA_CHECK; compute-sanitizer is not available on the GPU pool): every raw
// shared-memory address the consumers form, every cell id, label and row index is tested against its bounds; a violation prints
// where and traps, so the call fails loudly.  The GPU test suite runs against such a build with CLUSTASSESS_B200_LIB.
#ifdef CA_CHECK
#define CA_CHK(cond, what)                                                                                                          \
    do {                                                                                                                            \
        if (!(cond)) {                                                                                                              \
            printf("CA_CHECK failed: %s (%s:%d) block %d thread %d\n", what, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
            __trap();                                                                                                               \
        }                                                                                                                           \
    } while (0)
#else
#define CA_CHK(cond, what) \
    do {                   \
    } while (0)
#endif
static constexpr unsigned FULL = 0xffffffffu;
static constexpr int NC_RES = TL_THREADS;      // consumer threads of the resident kernel (a quad of cells each)
static constexpr int NC_STR = TL_STR_THREADS;  // consumer threads of the streaming kernel
static constexpr int CAP = TL_CAP;
static_assert(TL_CPT == 4, "a resident thread owns one quad of cells");
static_assert(NC_RES % 32 == 0 && NC_STR % 32 == 0 && NC_RES + 32 <= 1024 && NC_STR + 32 <= 1024, "whole warps, one block");

// shared-memory map (byte offsets from the dynamic shared base)
static constexpr int SM_MBAR = 0;       // four mbarriers: full[2] (tables initialised + meta published), empty[2] (buffer released)
static constexpr int SM_META = 32;      // two phase records
static constexpr int META_BYTES = 224;
// phase record.  Header word: cnt | g0 << 8 | first << 16 | big << 17 | WU-record parity << 18 (read together with j0 by one
// 8-byte load); everything per partner is indexed by TILE SLOT (partner j0 + slot), so the consumers' unrolled slot bodies
// address it with immediates
static constexpr int META_HDR = 0, META_J0 = 4,
                     META_TR = 64,   // [8] {table base, row bytes}
                     META_W = 160;   // [8] double: weights (not written when every weight is 1)
static constexpr int SM_SIMACC = 480;   // [2][8] u64: fixed-point sum of ECS over the block's cells, per buffer and tile slot
static constexpr int SM_WUREC = 608;    // two work-unit records (alternating), written once per work unit
// work-unit record: the item, and where the NEXT work unit of this block starts (0 positions = none), for the prefetch
static constexpr int WU_M = 0, WU_POS0 = 4, WU_NPOS = 8, WU_NROWS = 12, WU_SA = 16, WU_SLOT = 20, WU_JHI = 24, WU_NXT_ST = 28,
                     WU_NXT_SLOT = 32, WU_NXT_POS0 = 36, WU_NXT_NPOS = 40, WUREC_BYTES = 48;
static constexpr int SM_PLAN = 704;     // producer's scratch for the current span: [32] {row bytes, inclusive prefix of the row
static constexpr int SM_PLANW = 960;    //   bytes inside the 8-partner tile}; [32] weights
static constexpr int SM_DYN = 1280;     // two table buffers
static_assert(SM_WUREC + 2 * WUREC_BYTES <= SM_PLAN && SM_PLAN + 256 <= SM_PLANW && SM_PLANW + 256 <= SM_DYN, "shared-memory map");
static constexpr uint32_t KK_BIG = 1u << 30;  // partner has a cluster with >= 65535 cells
// the instantiations of the kernel: per-cell everything; ratio tables (TileParams::rt); the matrix from the table entries (TileParams::sim_scan)
static constexpr int MODE_PLAIN = 0, MODE_RT = 1, MODE_SCAN = 2;
// ratio-table launches (RT): bytes of a table buffer the count tables may use; the ratio tables (8 bytes per entry) follow them
__host__ __device__ __forceinline__ int rt_count_bytes(int half) { return (half / 3) & ~127; }

// ---- shared-space accessors ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
// the immediate 1 matters: ptxas emits ATOMS.POPC.INC (same-address lanes merged in hardware) only for it
__device__ __forceinline__ void red_inc(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }
__device__ __forceinline__ void red_add_u64(uint32_t a, unsigned long long v) { asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long lds_b64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_b64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
// double -> fixed point (round to nearest); the product is far below 2^63 by construction
__device__ __forceinline__ unsigned long long to_fix(double v, double scale) { return (unsigned long long)__double2ll_rn(v * scale); }

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
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t spin = 1; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (BACKOFF && CA_BACKOFF_NS > 0 && !done) __nanosleep(CA_BACKOFF_NS);  // the producer is usually far ahead: leave the issue slots to the consumers
        if ((spin & 0xfffffu) == 0) {  // every ~0.1 s: a lost copy or arrive must fail loudly, never hang the GPU
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            if (t - t0 > 120ull * 1000000000ull) __trap();  // 2 minutes: far beyond any legitimate phase
        }
    }
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ uint4 ldg_u4(const uint16_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// exact for every uint32: 2^52 + v is representable and the subtraction is exact
__device__ __forceinline__ double u32_to_f64(uint32_t v) { return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0; }

// labels of a whole super-tile (16 partners) for a quad of cells: ONE 32-byte load per cell (SASS LDG.E.256, sm_100+),
// i.e. one L2 sector request per (cell, work unit); lo = partners [16 s, 16 s + 8), hi = [16 s + 8, 16 s + 16)
__device__ __forceinline__ void ldg_256(const uint16_t* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}
// padding positions (cell id -1) read the pad cell n, whose label in partition p is K_p: the spare table column
__device__ __forceinline__ int64_t cell_or_pad(int cell, uint32_t n) { return (int64_t)min((uint32_t)cell, n); }
__device__ __forceinline__ void load_quad_super(const uint16_t* __restrict__ lab16s, uint32_t n, int4 cell, uint4 (&Lo)[4], uint4 (&Hi)[4]) {
    ldg_256(lab16s + cell_or_pad(cell.x, n) * 16, Lo[0], Hi[0]);
    ldg_256(lab16s + cell_or_pad(cell.y, n) * 16, Lo[1], Hi[1]);
    ldg_256(lab16s + cell_or_pad(cell.z, n) * 16, Lo[2], Hi[2]);
    ldg_256(lab16s + cell_or_pad(cell.w, n) * 16, Lo[3], Hi[3]);
}
// labels of the 8 partners [j0, j0+8) for a quad of cells: one 16-byte load per cell (lab16j = lab16 + lab16_index(n, 0, j0))
__device__ __forceinline__ void load_quad(const uint16_t* __restrict__ lab16j, uint32_t n, int4 cell, uint4 (&L)[4]) {
    L[0] = ldg_u4(lab16j + cell_or_pad(cell.x, n) * 16);
    L[1] = ldg_u4(lab16j + cell_or_pad(cell.y, n) * 16);
    L[2] = ldg_u4(lab16j + cell_or_pad(cell.z, n) * 16);
    L[3] = ldg_u4(lab16j + cell_or_pad(cell.w, n) * 16);
}

struct Phase {
    int j0, g0, cnt;  // partners [j0+g0, j0+g0+cnt) of the 8-partner tile at j0
};

// ---- histogram / gather of one phase ---------------------------------------------------------------------------
// A phase covers the tile slots [g0, g0 + cnt) of the 8-partner tile whose labels sit in L.  The slot bodies are
// instantiated per slot (the label's word and half are compile-time, the slot's meta entries are at immediate
// offsets) and entered Duff-style: one jump to slot g0, then one compare-and-branch per slot.  There are no
// validity tests: padding positions carry the pad label (ca_internal.h, lab16_index) and count into a spare column.
// A thread's cells come in GROUPS of 4 (a resident thread has one group, a streaming thread up to SCPT / 4; `ng` =
// groups in use, block-uniform).
template <int V> struct Slot { static constexpr int value = V; };
template <typename F>
__device__ __forceinline__ void for_slots(int g0, int cnt, F&& f) {
    int left = cnt;
    switch (g0) {
        case 0: f(Slot<0>{}); if (--left == 0) break; [[fallthrough]];
        case 1: f(Slot<1>{}); if (--left == 0) break; [[fallthrough]];
        case 2: f(Slot<2>{}); if (--left == 0) break; [[fallthrough]];
        case 3: f(Slot<3>{}); if (--left == 0) break; [[fallthrough]];
        case 4: f(Slot<4>{}); if (--left == 0) break; [[fallthrough]];
        case 5: f(Slot<5>{}); if (--left == 0) break; [[fallthrough]];
        case 6: f(Slot<6>{}); if (--left == 0) break; [[fallthrough]];
        default: f(Slot<7>{}); break;
    }
}
// sel = the PRMT selector of the label's half (P.sel[0] = 0x4410 low, P.sel[1] = 0x4432 high).  It is read from the
// kernel parameters on purpose: with an immediate selector ptxas rewrites the extraction + address into three
// instructions (shift, mask, add); with an opaque one it stays PRMT + IMAD.
template <int JT>
__device__ __forceinline__ uint32_t slot_label(const TileParams& P, const uint4& l) {
    const uint32_t w = (JT >> 1) == 0 ? l.x : (JT >> 1) == 1 ? l.y : (JT >> 1) == 2 ? l.z : l.w;
    return __byte_perm(w, 0u, P.sel[JT & 1]);
}
__device__ __forceinline__ uint32_t slot_label_rt(const uint4& l, int jt) {
    const int wi = jt >> 1;
    const uint32_t w = wi == 0 ? l.x : wi == 1 ? l.y : wi == 2 ? l.z : l.w;
    return (jt & 1) ? (w >> 16) : (w & 0xffffu);
}

// address of entry `lab` of the table row at rb (row bytes rowb, entry bytes esz); checked builds test it
__device__ __forceinline__ uint32_t entry_addr(const TileParams& P, uint32_t rb, uint32_t rowb, uint32_t lab, uint32_t esz) {
#ifdef CA_CHECK
    extern __shared__ __align__(128) unsigned char smem_chk[];
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_chk);
    CA_CHK(lab * 4u + 4u <= rowb, "label beyond the table row");
    CA_CHK(rb + lab * esz >= sb + SM_DYN && rb + lab * esz + esz <= sb + (uint32_t)P.smem_bytes, "table entry outside the table buffers");
#endif
    return rb + lab * esz;
}

template <int NCELL>
__device__ __forceinline__ void hist_cells(const TileParams& P, uint32_t meta, const Phase ph, const uint4 (&L)[NCELL], uint32_t rowidx, int ng) {
    for_slots(ph.g0, ph.cnt, [&](auto slot) {
        constexpr int JT = decltype(slot)::value;
        const uint2 tr = lds_u64(meta + META_TR + JT * 8);
        const uint32_t rb = tr.x + rowidx * tr.y;
#pragma unroll
        for (int g = 0; g < NCELL / 4; g++) {
            if (NCELL == 4 || g < ng) {
                red_inc(entry_addr(P, rb, tr.y, slot_label<JT>(P, L[4 * g + 0]), 4u));
                red_inc(entry_addr(P, rb, tr.y, slot_label<JT>(P, L[4 * g + 1]), 4u));
                red_inc(entry_addr(P, rb, tr.y, slot_label<JT>(P, L[4 * g + 2]), 4u));
                red_inc(entry_addr(P, rb, tr.y, slot_label<JT>(P, L[4 * g + 3]), 4u));
            }
        }
    });
}

// ECS of one (cell, partner): T[a][b] / max(S_i[a], S_j[b])  (calculate_ecs.cpp:35) as count * r, r = 1/max from
// rcp.approx.ftz.f64 + one Newton step (relative error < 1e-13).  uint32 -> double goes through the 2^52 trick (one
// DADD on the fp64 pipe) instead of I2F.F64: with two I2F and the MUFU per unit the gather is bound by the
// quarter-rate XU pipe (ncu: 85 % busy).
__device__ __forceinline__ double ecs_term(uint32_t count, uint32_t sa, uint32_t sb) {
    const double mx = u32_to_f64(max(sa, sb));
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(mx));
    r = fma(fma(-mx, r, 1.0), r, r);
    return u32_to_f64(count) * r;
}

// The matrix entry of a pair is sum_n ECS_ij[n] / N.  To make it independent of which thread holds which cell (the order of
// the cells inside a cluster varies from run to run, ca_pack.cu) every term is rounded to the fixed-point grid 2^-F BEFORE
// it is added: (v + magic) - magic with magic = 1.5 * 2^(52 - F).  Sums of grid values below 2^(53 - F) are EXACT in fp64
// (F <= 40, a warp adds at most 128 terms <= 1), so a thread's and a warp's partial sums do not depend on the order of the
// additions; the warp's total then joins the block's 64-bit integer accumulator of its tile slot, and integers are associative.
__device__ __forceinline__ double sim_round(const TileParams& P, double v) { return (v + P.sim_magic) - P.sim_magic; }
// A thread's sum `es` is an exact multiple of 2^-F below 2^11: es + magic carries es * 2^F as an integer in its low mantissa bits
// (magic's own low word is zero), so the warp total is two hardware integer reductions (REDUX) of its 24-bit pieces instead of a
// shuffle tree of doubles.
__device__ __forceinline__ void sim_warp_add(const TileParams& P, uint32_t simacc_slot, double es, int lane) {
    const double t = es + P.sim_magic;
    const uint32_t lo = (uint32_t)__double2loint(t), hi = (uint32_t)(__double2hiint(t) - __double2hiint(P.sim_magic));
    const uint32_t s0 = __reduce_add_sync(FULL, lo & 0xffffffu);
    const uint32_t s1 = __reduce_add_sync(FULL, __funnelshift_r(lo, hi, 24));
    if (lane == 0 && (s0 | s1) != 0u) red_add_u64(simacc_slot, ((unsigned long long)s1 << 24) + s0);
}

// UNITW: all weights are 1 (element_consistency of distinct partitions): ECS = count * r is accumulated by one FMA.
// BIG (block-uniform, rare): a partner of the phase has a cluster with >= 65535 cells, whose size half reads 0xffff
// = "look the size up in global memory"; WIDE: a streaming item with more than 65535 cells, 32-bit counters and all
// partner sizes from global memory.  Both take the generic loop with run-time slots.
// vmask: bit k = cell k is a real cell (padding positions gather the spare column; they must not count in the matrix).
template <int NCELL, bool WIDE, bool WANT_SIM, bool UNITW>
__device__ __forceinline__ void gather_cells(const TileParams& P, uint32_t meta, uint32_t simacc, const Phase ph, bool big,
                                             const uint4 (&L)[NCELL], uint32_t rowidx, uint32_t sa, uint32_t vmask, int ng, int lane,
                                             double (&acc)[NCELL]) {
    if (!WIDE && !big) {
        for_slots(ph.g0, ph.cnt, [&](auto slot) {
            constexpr int JT = decltype(slot)::value;
            const uint2 tr = lds_u64(meta + META_TR + JT * 8);
            const uint32_t rb = tr.x + rowidx * tr.y;
            const double wj = UNITW ? 1.0 : lds_f64(meta + META_W + JT * 8);
            double es = 0.0;
#pragma unroll
            for (int g = 0; g < NCELL / 4; g++) {
                if (NCELL == 4 || g < ng) {
#pragma unroll
                    for (int c = 4 * g; c < 4 * g + 4; c++) {
                        const uint32_t e = lds_u32(entry_addr(P, rb, tr.y, slot_label<JT>(P, L[c]), 4u));
                        const uint32_t count = e & 0xffffu, sb = e >> 16;
                        if (UNITW && !WANT_SIM) {
                            const double mx = u32_to_f64(max(sa, sb));
                            double r;
                            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(mx));
                            r = fma(fma(-mx, r, 1.0), r, r);
                            acc[c] = fma(u32_to_f64(count), r, acc[c]);
                        } else {
                            const double v = ecs_term(count, sa, sb);
                            acc[c] = UNITW ? acc[c] + v : fma(v, wj, acc[c]);
                            if (WANT_SIM) es += ((vmask >> c) & 1u) ? sim_round(P, v) : 0.0;
                        }
                    }
                }
            }
            if (WANT_SIM) sim_warp_add(P, simacc + JT * 8u, es, lane);
        });
        return;
    }
#pragma unroll 1
    for (int jt = ph.g0; jt < ph.g0 + ph.cnt; jt++) {
        const uint2 tr = lds_u64(meta + META_TR + jt * 8);
        const uint32_t rb = tr.x + rowidx * tr.y;
        const double wj = UNITW ? 1.0 : lds_f64(meta + META_W + jt * 8);
        const int32_t* szg = P.sizes + (int64_t)(ph.j0 + jt) * P.kp;
        double es = 0.0;
#pragma unroll
        for (int g = 0; g < NCELL / 4; g++) {
            if (NCELL == 4 || g < ng) {
#pragma unroll
                for (int c = 4 * g; c < 4 * g + 4; c++) {
                    const uint32_t lab = slot_label_rt(L[c], jt);
                    const uint32_t e = lds_u32(entry_addr(P, rb, tr.y, lab, 4u));
                    uint32_t count, sb;
                    if (WIDE) {
                        count = e;
                        sb = (uint32_t)__ldg(szg + lab);
                    } else {
                        count = e & 0xffffu;
                        sb = e >> 16;
                        if (sb == 0xffffu) sb = (uint32_t)__ldg(szg + lab);
                    }
                    const double v = ecs_term(count, sa, sb);
                    acc[c] = UNITW ? acc[c] + v : fma(v, wj, acc[c]);
                    if (WANT_SIM) es += ((vmask >> c) & 1u) ? sim_round(P, v) : 0.0;
                }
            }
        }
        if (WANT_SIM) sim_warp_add(P, simacc + jt * 8u, es, lane);
    }
}

// the per-partner ECS sums of a finished phase go to the ECS matrix (called by 8 threads after a block barrier)
__device__ __forceinline__ void flush_sim(const TileParams& P, uint32_t simacc, const Phase ph, int m, int t) {
    if (m < 0) return;
    if (t >= ph.g0 && t < ph.g0 + ph.cnt) {  // t = tile slot
        const unsigned long long v = lds_b64(simacc + t * 8u);
        CA_CHK(m < P.m && ph.j0 + t < P.m, "matrix entry outside [0, m) x [0, m)");
        if (v != 0ull) atomicAdd(P.sim + (int64_t)m * P.m + (ph.j0 + t), v);
        sts_b64(simacc + t * 8u, 0ull);
    }
}
// ---- the ECS matrix from the TABLES (MODE_SCAN launches: partitions with few clusters) -----------------------------------------------
// sum_n ECS_ij[n] = sum_{a,b} T[a][b]^2 / max(S_i[a], S_j[b]) -- the reference's own formula for the average of a pair
// (disjointECSaverage, src/calculate_ecs.cpp:58-67): every non-empty table entry stands for T[a][b] cells with the same ratio.
// With K clusters a table row has K entries but n / K cells, so for small K reading the finished tables is far cheaper than
// rounding and reducing one term per (cell, partner) in the gather (element_sim_matrix then needs no gather at all).  The
// consumer warps share the (partner, row) pairs of the phase; a warp's lanes stride over the row's columns.  The row's own
// cluster size is the sum of its counts (every cell of the cluster sits in exactly one real column; padding positions sit in
// the pad column K_j, which is not read).  Every term is converted to fixed point on its own and the integers are added, so
// the result does not depend on which warp or block reads which row.  WIDE: 32-bit counters, partner sizes from global memory.
template <bool WIDE>
__device__ __forceinline__ void scan_sim(const TileParams& P, uint32_t meta, uint32_t simacc, const Phase ph, int R, uint32_t sa_item,
                                         int wc, int nw, int lane) {
    const uint32_t kk = lane < 8 ? (uint32_t)__ldg(P.kinfo + ph.j0 + lane) : 0u;  // K_j | flags of the tile's partners
    const int npairs = ph.cnt * R;
#pragma unroll 1
    for (int q = wc; q < npairs; q += nw) {  // warp-uniform
        const int sl = q / R, row = q - sl * R, jt = ph.g0 + sl;
        const uint32_t K = __shfl_sync(FULL, kk, jt) & 0xffffffu;
        const uint2 tr = lds_u64(meta + META_TR + jt * 8);
        const uint32_t rb = tr.x + (uint32_t)row * tr.y;
        uint32_t sa = sa_item;
        if (sa == 0u) {
            uint32_t cs = 0u;
            for (uint32_t col = (uint32_t)lane; col < K; col += 32u) {
                const uint32_t e = lds_u32(rb + col * 4u);
                cs += WIDE ? e : (e & 0xffffu);
            }
            sa = __reduce_add_sync(FULL, cs);
            if (sa == 0u) continue;  // warp-uniform
        }
        const int32_t* szg = P.sizes + (int64_t)(ph.j0 + jt) * P.kp;
        unsigned long long sum = 0ull;
        for (uint32_t col = (uint32_t)lane; col < K; col += 32u) {
            const uint32_t e = lds_u32(rb + col * 4u);
            const uint32_t count = WIDE ? e : (e & 0xffffu);
            if (count != 0u) {
                uint32_t sb = WIDE ? 0xffffu : (e >> 16);
                if (sb == 0xffffu) sb = (uint32_t)__ldg(szg + col);
                const double mx = u32_to_f64(max(sa, sb));
                double r;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(mx));
                r = fma(fma(-mx, r, 1.0), r, r);
                const double c = u32_to_f64(count);
                sum += to_fix((c * c) * r, P.sim_scale);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        if (lane == 0 && sum != 0ull) red_add_u64(simacc + (uint32_t)jt * 8u, sum);
    }
}

// ---- ratio tables (RT kernels: partitions with few clusters) ------------------------------------------------------------------
// With K clusters a table has R x K entries for the item's ~n R / K cells, so for small K the ratio T[a][b] / max(S_a, S_b) is
// computed once per ENTRY instead of once per (cell, partner): after the histogram the consumer warps share the (partner, row)
// pairs of the phase and write, next to every count table, a table of doubles holding the finished ratio (IEEE division: each
// term is bit-identical to the reference's (double) T / max, src/calculate_ecs.cpp:35); after a second barrier the gather is one
// 8-byte shared load and one add per (cell, partner) instead of a load and 14 arithmetic instructions.  The count tables get a
// third of the buffer, the ratio tables the rest: ratio entry address = rbase + 2 * (count entry address - cbase).  The pad
// column gets 0 (padding positions read it).  The ECS matrix comes from the same pass (sum of T * ratio over the entries).
template <bool WIDE, bool WANT_SIM>
__device__ __forceinline__ void convert_tables(const TileParams& P, uint32_t meta, uint32_t simacc, const Phase ph, int R, uint32_t sa_item,
                                               uint32_t cbase, uint32_t rbase, int wc, int nw, int lane) {
    const uint32_t kk = lane < 8 ? (uint32_t)__ldg(P.kinfo + ph.j0 + lane) : 0u;  // K_j | flags of the tile's partners
    const int npairs = ph.cnt * R;
#pragma unroll 1
    for (int q = wc; q < npairs; q += nw) {  // warp-uniform
        const int sl = q / R, row = q - sl * R, jt = ph.g0 + sl;
        const uint32_t K = __shfl_sync(FULL, kk, jt) & 0xffffffu;
        const uint2 tr = lds_u64(meta + META_TR + jt * 8);
        const uint32_t rb = tr.x + (uint32_t)row * tr.y;
        const uint32_t rr = rbase + (rb - cbase) * 2u;
        uint32_t sa = sa_item;
        if (sa == 0u) {  // the row's cluster size = the sum of its counts
            uint32_t cs = 0u;
            for (uint32_t col = (uint32_t)lane; col < K; col += 32u) {
                const uint32_t e = lds_u32(rb + col * 4u);
                cs += WIDE ? e : (e & 0xffffu);
            }
            sa = __reduce_add_sync(FULL, cs);
        }
        const int32_t* szg = P.sizes + (int64_t)(ph.j0 + jt) * P.kp;
        unsigned long long sum = 0ull;
        for (uint32_t col = (uint32_t)lane; col <= K; col += 32u) {
            const uint32_t e = lds_u32(rb + col * 4u);
            const uint32_t count = WIDE ? e : (e & 0xffffu);
            double v = 0.0;
            if (count != 0u && col < K) {
                uint32_t sb = WIDE ? 0xffffu : (e >> 16);
                if (sb == 0xffffu) sb = (uint32_t)__ldg(szg + col);
                const double c = u32_to_f64(count);
                v = __ddiv_rn(c, u32_to_f64(max(sa, sb)));
                if (WANT_SIM) sum += to_fix(c * v, P.sim_scale);
            }
            sts_f64(rr + col * 8u, v);
        }
        if (WANT_SIM) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
            if (lane == 0 && sum != 0ull) red_add_u64(simacc + (uint32_t)jt * 8u, sum);
        }
    }
}
template <bool UNITW>
__device__ __forceinline__ void gather_ratio(const TileParams& P, uint32_t meta, const Phase ph, const uint4 (&L)[4], uint32_t rowidx,
                                             uint32_t cbase, uint32_t rbase, double (&acc)[4]) {
    for_slots(ph.g0, ph.cnt, [&](auto slot) {
        constexpr int JT = decltype(slot)::value;
        const uint2 tr = lds_u64(meta + META_TR + JT * 8);
        const uint32_t rr = rbase + (tr.x - cbase + rowidx * tr.y) * 2u;
        const double wj = UNITW ? 1.0 : lds_f64(meta + META_W + JT * 8);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double v = lds_f64(entry_addr(P, rr, tr.y, slot_label<JT>(P, L[c]), 8u));
            acc[c] = UNITW ? acc[c] + v : fma(v, wj, acc[c]);
        }
    });
}

// a cell's sum over the partners of a work unit (already scaled by w_i / norm) joins its fixed-point accumulator
__device__ __forceinline__ void ecc_add(const TileParams& P, int cell, double v) {
    CA_CHK(cell >= 0 && (int64_t)cell < P.n, "cell id outside [0, n)");
    atomicAdd(P.ecc + cell, to_fix(v, ECC_FIX_SCALE));
}

// ---- work units and phase planning (warp 0) -------------------------------------------------------------------
// Work unit (WU) = (item, span of WSPAN partners).  WUs are numbered span-major: WU w belongs to the
// span s with wu_prefix[s] <= w < wu_prefix[s+1] and to item w - wu_prefix[s]; the items are sorted by their
// first partner, so the items that have partners inside span s are a PREFIX of the item list.  Block b of
// the persistent grid starts with WU b and then takes the next WU nobody has taken yet (advance_wu): all SMs sweep the partner axis together, so the label
// sectors of the current super-tile (32 bytes per cell) are fetched from HBM once and then served by L2 to every
// owner partition -- instead of one random HBM sector read per (cell, tile, owner), which is what bounds a
// block-per-item schedule (measured: 42 G random sectors/s, tools/gatherbench.cu).
static constexpr int SUPER = TL_SUPER;
static constexpr int WSPAN = TL_WU_SPAN;  // partners per work unit (one or two super-tiles)
#ifdef CA_ABLATE
#define CA_DBG(P, bit) (((P).dbg >> (bit)) & 1)  // tuning builds: 0 no ecc flush, 1 no table fill, 2 no label loads, 3 no hist, 4 no gather, 5 no barrier between hist and gather
#else
#define CA_DBG(P, bit) 0
#endif
static constexpr uint32_t NROWS_WIDE = 1u << 30;  // Item::nrows flag: more than 65535 cells, 32-bit counters


struct Planner {          // the producer warp's registers (all values warp-uniform)
    int w;                // current WU
    int s;                // its span
    int wend;             // wu_prefix[s + 1]
    Item it;              // its item
    int jhi;              // partners of this WU: [.., jhi)
    int j0, g0;           // cursor: tile and first tile slot of the next phase to plan
    int first;            // the next phase is the first of its WU
    unsigned bigmask;     // partners of the current span that have a cluster with >= 65535 cells (bit = position in the span)
    int wpar;             // parity of the WU record the current WU uses
    Item nit;             // item of the NEXT WU of this block (loaded one WU ahead), nw/ns/nwend its ids
    int nw, ns, nwend;
};

// ids of WU w (w ascending over calls): advance the span while w is beyond it
__device__ __forceinline__ void locate_wu(const TileParams& P, int w, int& s, int& wend) {
    while (w >= wend && s + 1 < P.nsuper) {
        s++;
        wend = __ldg(P.wu_prefix + s + 1);
    }
}
__device__ __forceinline__ Item load_item(const TileParams& P, int w, int s) {
    const int4* q = reinterpret_cast<const int4*>(P.items + (w - __ldg(P.wu_prefix + s)));
    const int4 a = __ldg(q), b = __ldg(q + 1);
    Item it;
    it.m = a.x; it.pos0 = a.y; it.npos = a.z; it.nrows = a.w;
    it.jbeg = b.x; it.jend = b.y; it.sa = b.z; it.pad_ = b.w;
    return it;
}
// make the prefetched WU current (or mark the end) and start loading the one after it
__device__ __forceinline__ void advance_wu(const TileParams& P, Planner& pl, int lane, uint32_t sbase) {
    const int olds = pl.w < 0 ? -1 : pl.s;
    pl.w = pl.nw;
    pl.s = pl.ns;
    pl.wend = pl.nwend;
    pl.it = pl.nit;
    if (pl.w < P.nwu) {
        const int slo = pl.s * WSPAN;
        const int jlo = max(pl.it.jbeg, slo);
        pl.jhi = min(pl.it.jend, slo + WSPAN);
        pl.j0 = jlo & ~7;
        pl.g0 = jlo - pl.j0;
        pl.first = 1;
        pl.wpar ^= 1;
        if (pl.s != olds) {  // a new span: row bytes of its partners' tables, their prefix sums inside each 8-partner tile, flags, weights
            const uint32_t kk = lane < WSPAN ? (uint32_t)__ldg(P.kinfo + slo + lane) : 0u;
            const uint32_t rowb = (((kk & 0xffffu) + 1u + 3u) & ~3u) * 4u;  // bytes of one table row: K_j entries + the pad column
            uint32_t pre = rowb;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, pre, o);
                if ((lane & 7) >= o) pre += v;
            }
            pl.bigmask = __ballot_sync(FULL, (kk & KK_BIG) != 0);
            __syncwarp();  // every lane has finished reading the previous span's entries (plan_compute of the last phase)
            if (lane < WSPAN) {
                sts_u32(sbase + SM_PLAN + lane * 8, rowb);
                sts_u32(sbase + SM_PLAN + lane * 8 + 4, pre);
                sts_f64(sbase + SM_PLANW + lane * 8, __ldg(P.w + slo + lane));
            }
            __syncwarp();
        }
        // the consumers read this WU's cell ids and quad info (streamed from HBM, 6 bytes per position) when they
        // reach it, one or two phases from now (or a tile earlier, with the prefetch): ask L2 for the lines already
        if (pl.it.npos <= CAP) {
            const int64_t off = (int64_t)pl.it.pad_ * P.perm_stride + pl.it.pos0;
            const char* pc = reinterpret_cast<const char*>(P.perm + off);
            const char* pq = reinterpret_cast<const char*>(reinterpret_cast<const uint2*>(P.qinfo) + (off >> 2));
            for (int b = lane * 128; b < pl.it.npos * 4; b += 32 * 128) prefetch_l2(pc + b);
            for (int b = lane * 128; b < pl.it.npos * 2; b += 32 * 128) prefetch_l2(pq + b);
        }
        // This block's next work unit: the first one nobody has taken yet.  Work units differ a lot in length (a streaming item
        // has 3.5k .. 100k+ cells, a resident one 1 .. hundreds of table rows), so dealing them round-robin left blocks idle at
        // the end of a launch (measured at config 5: the eight per-device streaming launches of an 8-GPU job summed to 46 ms
        // against 35.5 ms on one GPU).  The order of the hand-out is still the span-major order, so the sweep over the label
        // slabs stays together; sums are integers, so who computes what does not change a bit of the result.
        if (P.wu_counter) {
            int nxt = 0;
            if (lane == 0) nxt = (int)gridDim.x + atomicAdd(P.wu_counter, 1);
            pl.nw = __shfl_sync(FULL, nxt, 0);
        } else {
            pl.nw = pl.w + (int)gridDim.x;
        }
        if (pl.nw < P.nwu) {
            locate_wu(P, pl.nw, pl.ns, pl.nwend);
            pl.nit = load_item(P, pl.nw, pl.ns);
        }
    }
}
__device__ __forceinline__ void planner_init(const TileParams& P, Planner& pl, int lane, uint32_t sbase) {
    pl.w = -1;
    pl.wpar = 0;
    pl.bigmask = 0;
    pl.nw = (int)blockIdx.x;
    pl.ns = 0;
    pl.nwend = __ldg(P.wu_prefix + 1);
    if (pl.nw < P.nwu) {
        locate_wu(P, pl.nw, pl.ns, pl.nwend);
        pl.nit = load_item(P, pl.nw, pl.ns);
    }
    advance_wu(P, pl, lane, sbase);
}

// Plans the phase at the cursor (how many partners of the tile fit into one buffer: R rows of K_j entries each,
// K_j + 1 rounded to 4) -- this part needs no buffer, so the producer does it BEFORE it waits for the buffer to be
// released -- then publishes it into buffer `buf`: table addresses, the WU's item, and the bulk copies that
// initialise the tables.  After the last WU it publishes cnt = 0.  Everything written here is published by the
// arrive on the buffer's `full` mbarrier.
struct PhasePlan {
    uint32_t rowb, toff, ttot, sq;  // per lane (= tile slot g0 + lane): row bytes, table offset, position in the span; total bytes
    int cnt, R;
    bool wide, big;
};
__device__ __forceinline__ PhasePlan plan_compute(const TileParams& P, uint32_t sbase, int half, int lane, const Planner& pl) {
    PhasePlan pp;
    pp.wide = (pl.it.nrows & NROWS_WIDE) != 0;
    pp.R = pl.it.npos <= CAP ? (int)(pl.it.nrows & 0xffff) : 1;
    const int thi = min(8, pl.jhi - pl.j0);
    const int q = pl.g0 + lane;                                   // tile slot served by this lane
    const uint32_t sq0 = ((uint32_t)pl.j0 & (WSPAN - 1)) + (uint32_t)pl.g0;  // position of the phase's first partner inside the span
    pp.sq = (sq0 + (uint32_t)lane) & (WSPAN - 1);
    const uint2 rp = lds_u64(sbase + SM_PLAN + pp.sq * 8);        // {row bytes, inclusive prefix inside the tile}
    const uint32_t base = pl.g0 ? lds_u32(sbase + SM_PLAN + (sq0 - 1u) * 8 + 4) : 0u;
    const bool ok = lane < 8 && q < thi;
    pp.rowb = ok ? rp.x : 0u;
    const uint32_t pre = (rp.y - base) * (uint32_t)pp.R;          // bytes of the tables of slots g0 .. g0 + lane
    const unsigned fm = __ballot_sync(FULL, ok && pre <= (uint32_t)half);  // monotone: the set bits are a prefix
    pp.cnt = __popc(fm & 0xffu);                                  // >= 1: the host bounds R so that one partner always fits
    if (pp.cnt == 0) __trap();                                    // cannot happen; never hang the GPU
    pp.ttot = __shfl_sync(FULL, pre, pp.cnt - 1);
    pp.toff = pre - pp.rowb * (uint32_t)pp.R;
    pp.big = ((pl.bigmask >> sq0) & ((1u << pp.cnt) - 1u)) != 0;
    CA_CHK(pp.ttot <= (uint32_t)half && pp.toff + pp.rowb * (uint32_t)pp.R <= (uint32_t)half || lane >= pp.cnt, "phase tables beyond the buffer");
    CA_CHK(pl.it.m >= 0 && pl.it.m < P.m && pl.j0 + pl.g0 + pp.cnt <= P.m, "work unit outside the partition range");
    CA_CHK((int64_t)pl.it.pos0 + pl.it.npos <= P.perm_stride && pl.it.pos0 >= 0 && pl.it.npos > 0, "item outside its partition's sorted order");
    return pp;
}
template <bool UNITW>
__device__ __forceinline__ void plan_publish(const TileParams& P, uint32_t sbase, uint32_t buf, uint32_t bufaddr, int lane,
                                             const Planner& pl, const PhasePlan& pp) {
    const uint32_t meta = sbase + SM_META + buf * META_BYTES, bar = sbase + SM_MBAR + buf * 8u;
    const uint32_t tb = bufaddr + pp.toff;
    if (lane < pp.cnt) {
        const uint32_t slot = (uint32_t)(pl.g0 + lane);
        sts_u64(meta + META_TR + slot * 8, tb, pp.rowb);
        if (!UNITW) sts_f64(meta + META_W + slot * 8, lds_f64(sbase + SM_PLANW + pp.sq * 8));
    }
    if (lane == 0) {
        sts_u64(meta + META_HDR,
                (uint32_t)pp.cnt | ((uint32_t)pl.g0 << 8) | ((uint32_t)pl.first << 16) | (pp.big ? 1u << 17 : 0u) | ((uint32_t)pl.wpar << 18),
                (uint32_t)pl.j0);
        if (pl.first) {  // the work unit's record: read by the consumers at its first phase (every phase by the streaming kernel)
            const uint32_t wr = sbase + SM_WUREC + (uint32_t)pl.wpar * WUREC_BYTES;
            sts_u64(wr + WU_M, (uint32_t)pl.it.m, (uint32_t)pl.it.pos0);
            sts_u64(wr + WU_NPOS, (uint32_t)pl.it.npos, (uint32_t)pl.it.nrows);
            sts_u64(wr + WU_SA, (uint32_t)pl.it.sa, (uint32_t)pl.it.pad_);
            // where this block's next work unit starts: its first super-tile, and its cells in perm
            uint32_t nst = 0u, nslot = 0u, npos0 = 0u, nnpos = 0u;
            if (pl.nw < P.nwu) {
                nst = (uint32_t)(max(pl.nit.jbeg, pl.ns * WSPAN) >> 4);
                nslot = (uint32_t)pl.nit.pad_;
                npos0 = (uint32_t)pl.nit.pos0;
                nnpos = (uint32_t)pl.nit.npos;
            }
            sts_u64(wr + WU_JHI, (uint32_t)pl.jhi, nst);
            sts_u64(wr + WU_NXT_SLOT, nslot, npos0);
            sts_u32(wr + WU_NXT_NPOS, nnpos);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the consumers' generic accesses before the async writes
    __syncwarp();
    if (CA_DBG(P, 1)) {
        if (lane == 0) mbar_arrive(bar);
        return;
    }
    // one bulk copy per partner: R rows of the partner's table template (ca_pack.cu, table_template_kernel).  The copy
    // instruction takes uniform operands, so the lanes' copies are issued one after the other: their number matters.
    if (lane == 0) {
        mbar_expect_tx(bar, pp.ttot);
        if (pp.wide) bulk_g2s(bufaddr, P.zeros, pp.ttot, bar);
    }
    if (!pp.wide && lane < pp.cnt)
        bulk_g2s(tb, reinterpret_cast<const char*>(P.tmpl) + (int64_t)(pl.j0 + pl.g0 + lane) * P.tmpl_bytes, pp.rowb * (uint32_t)pp.R, bar);
}
__device__ __forceinline__ void plan_terminal(uint32_t sbase, uint32_t buf, int lane) {
    if (lane == 0) {
        sts_u32(sbase + SM_META + buf * META_BYTES + META_HDR, 0u);
        mbar_arrive(sbase + SM_MBAR + buf * 8u);
    }
}
__device__ __forceinline__ void plan_advance(const TileParams& P, Planner& pl, const PhasePlan& pp, int lane, uint32_t sbase) {
    const int thi = min(8, pl.jhi - pl.j0);
    pl.first = 0;
    pl.g0 += pp.cnt;
    if (pl.g0 >= thi) {
        pl.j0 += 8;
        pl.g0 = 0;
        if (pl.j0 >= pl.jhi) advance_wu(P, pl, lane, sbase);
    }
}

template <bool UNITW>
__device__ __forceinline__ void producer_loop(const TileParams& P, uint32_t sbase, int half, int budget, int lane) {
    Planner pl;  // budget = bytes of a buffer the count tables may use (all of it, or a third when ratio tables follow them)
    planner_init(P, pl, lane, sbase);
    for (uint32_t p = 0;; p++) {
        const uint32_t b = p & 1u;
        const bool last = pl.w >= P.nwu;
        PhasePlan pp;
        if (!last) pp = plan_compute(P, sbase, budget, lane, pl);
        if (p >= 2) mbar_wait<true>(sbase + SM_MBAR + 16 + b * 8u, ((p >> 1) - 1u) & 1u);  // consumers are done with buffer b
        if (last) {
            plan_terminal(sbase, b, lane);
            break;
        }
        plan_publish<UNITW>(P, sbase, b, sbase + SM_DYN + b * (uint32_t)half, lane, pl, pp);
        plan_advance(P, pl, pp, lane, sbase);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The persistent kernel: one block per SM walks its work units; the phases of consecutive WUs form one stream.
// Warp 0 is the PRODUCER: it plans phases and issues their table fills, running up to two phases ahead of the
// CONSUMER warps (every other warp).  Hand-offs are mbarriers: full[b] (fill landed + meta published) and empty[b]
// (every consumer warp has finished gathering from buffer b); the consumers meet at one named barrier per phase,
// between histogram and gather.
// ---------------------------------------------------------------------------------------------------------------
template <int NCT>
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

struct PhaseHdr {
    Phase ph;
    bool first, big;
    uint32_t wurec;
};
__device__ __forceinline__ PhaseHdr read_hdr(uint32_t sbase, uint32_t meta) {
    const uint2 hdr = lds_u64(meta + META_HDR);  // {cnt | g0 << 8 | first << 16 | big << 17 | WU-record parity << 18, j0}
    PhaseHdr h;
    h.ph.cnt = (int)(hdr.x & 0xffu);
    h.ph.g0 = (int)((hdr.x >> 8) & 0xffu);
    h.ph.j0 = (int)hdr.y;
    h.first = (hdr.x >> 16) & 1u;
    h.big = (hdr.x >> 17) & 1u;
    h.wurec = sbase + SM_WUREC + ((hdr.x >> 18) & 1u) * WUREC_BYTES;
    return h;
}

// ---- resident items: a quad of consecutive positions per thread -------------------------------------------------
// The plain consumer: the 16 labels of a cell for a whole super-tile come with ONE 32-byte load when the work unit reaches
// the super-tile (L = its lower tile, Lh = its upper tile).
template <bool WANT_ECC, bool WANT_SIM, bool UNITW, int MODE>
__device__ __forceinline__ void consume_resident(const TileParams& P, const uint32_t sbase, const int tid, const int lane, const int warp) {
    constexpr bool RT = MODE == MODE_RT, SCAN = MODE == MODE_SCAN;
    const int half = ((P.smem_bytes - SM_DYN) / 2) & ~127;  // RT: a buffer = count tables (a third) + ratio tables
    const uint32_t third = (uint32_t)rt_count_bytes(half);
    int m = -1;
    int4 cell = make_int4(-1, -1, -1, -1);
    uint32_t vmask = 0;
    bool wactive = false;
    uint32_t rowidx = 0, sa = 0;
    int nrows = 1;          // WANT_SIM: table rows of the item
    uint32_t sa_item = 0;   // WANT_SIM: the cluster size of a single-row item (0 = several rows: scan_sim sums the row)
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    uint4 L[4], Lh[4];  // labels of the current tile for this thread's quad; Lh: the super-tile's upper tile
    int cur_j0 = -1;
    Phase prev;
    prev.j0 = prev.g0 = prev.cnt = 0;
    int prev_m = -1;

    auto flush_acc = [&]() {
        if (WANT_ECC && m >= 0 && !CA_DBG(P, 0)) {
            const double wscale = __ldg(P.w + m) * P.inv_norm;
            if (cell.x >= 0) ecc_add(P, cell.x, acc[0] * wscale);
            if (cell.y >= 0) ecc_add(P, cell.y, acc[1] * wscale);
            if (cell.z >= 0) ecc_add(P, cell.z, acc[2] * wscale);
            if (cell.w >= 0) ecc_add(P, cell.w, acc[3] * wscale);
        }
    };

    uint32_t p = 0;
    for (;; p++) {
        const uint32_t b = p & 1u;
        const uint32_t meta = sbase + SM_META + b * META_BYTES, simacc = sbase + SM_SIMACC + b * 64u;
        mbar_wait(sbase + SM_MBAR + b * 8u, (p >> 1) & 1u);  // buffer b initialised, meta of phase p published
        const uint2 hdr = lds_u64(meta + META_HDR);  // {cnt | g0 << 8 | first << 16 | big << 17 | WU-record parity << 18, j0}
        Phase ph;
        ph.cnt = (int)(hdr.x & 0xffu);
        if (ph.cnt == 0) break;
        ph.g0 = (int)((hdr.x >> 8) & 0xffu);
        ph.j0 = (int)hdr.y;
        const bool big = (hdr.x >> 17) & 1u;
        if ((hdr.x >> 16) & 1u) {  // a new work unit: retire the old one, pick up the new item
            flush_acc();
            const uint32_t wurec = sbase + SM_WUREC + ((hdr.x >> 18) & 1u) * WUREC_BYTES;
            const uint2 w0 = lds_u64(wurec + WU_M), w1 = lds_u64(wurec + WU_NPOS), w2 = lds_u64(wurec + WU_SA);
            m = (int)w0.x;
            const int npos = (int)w1.x;
            if (WANT_SIM || RT) {
                nrows = (int)(w1.y & 0xffffu);
                sa_item = nrows == 1 ? w2.x : 0u;
            }
            const int64_t off = (int64_t)w2.y * P.perm_stride + (int)w0.y;
            cur_j0 = -1;
            cell = make_int4(-1, -1, -1, -1);
            uint2 qi = make_uint2(0u, 0u);
            wactive = (tid & ~31) * 4 < npos;  // warp-uniform: the warp owns at least one position of the item
            if (tid * 4 < npos) {
                cell = __ldg(reinterpret_cast<const int4*>(P.perm + off + tid * 4));
                qi = __ldg(reinterpret_cast<const uint2*>(P.qinfo) + (off >> 2) + tid);
            }
            vmask = (1u << ((cell.x >= 0) + (cell.y >= 0) + (cell.z >= 0) + (cell.w >= 0))) - 1u;  // padding sits at the end of a cluster
            rowidx = qi.x;
            sa = qi.y;
            CA_CHK(min(min(cell.x, cell.y), min(cell.z, cell.w)) >= -1 && (int64_t)max(max(cell.x, cell.y), max(cell.z, cell.w)) < P.n, "cell id outside [-1, n)");
            CA_CHK(rowidx < (w1.y & 0xffffu) && (tid * 4 >= npos || sa > 0u), "quad info: row beyond the item's rows, or an empty cluster");
            acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
        }
        // ---- histogram ----
        if (wactive) {
            if (ph.j0 != cur_j0) {
                if (cur_j0 < 0 || ((ph.j0 ^ cur_j0) & ~(SUPER - 1)) != 0) {  // the WU's first phase or its second super-tile: L2 hits, the sectors were fetched by the first block that needed them
                    if (!CA_DBG(P, 2)) {
                        load_quad_super(P.lab16 + lab16_index(P.n, 0, ph.j0 & ~(SUPER - 1)), (uint32_t)P.n, cell, L, Lh);
                    } else {  // tuning builds: no label loads, every cell carries label 0
#pragma unroll
                        for (int c = 0; c < 4; c++) L[c] = Lh[c] = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                if ((ph.j0 & 8) != 0) {
#pragma unroll
                    for (int c = 0; c < 4; c++) L[c] = Lh[c];
                }
                cur_j0 = ph.j0;
            }
            if (!CA_DBG(P, 3)) hist_cells<4>(P, meta, ph, L, rowidx, 1);
        }
        if (!CA_DBG(P, 5)) consumer_sync<NC_RES>();  // the tables of phase p are complete; every consumer has left phase p-1
        if (WANT_SIM && warp == 1) flush_sim(P, sbase + SM_SIMACC + (b ^ 1u) * 64u, prev, prev_m, lane);
        // ---- gather ----
        if (RT) {  // ratio tables: one division per table entry, then one load and one add per (cell, partner)
            const uint32_t cbase = sbase + SM_DYN + b * (uint32_t)half, rbase = cbase + third;
            convert_tables<false, WANT_SIM>(P, meta, simacc, ph, nrows, sa_item, cbase, rbase, warp - 1, NC_RES / 32, lane);
            consumer_sync<NC_RES>();
            if (wactive) gather_ratio<UNITW>(P, meta, ph, L, rowidx, cbase, rbase, acc);
        } else if (WANT_SIM && SCAN) {  // the matrix from the tables; per-cell sums only when ecc is wanted
            scan_sim<false>(P, meta, simacc, ph, nrows, sa_item, warp - 1, NC_RES / 32, lane);
            if (WANT_ECC && wactive) gather_cells<4, false, false, UNITW>(P, meta, simacc, ph, big, L, rowidx, sa, vmask, 1, lane, acc);
        } else if (wactive && !CA_DBG(P, 4))
            gather_cells<4, false, WANT_SIM, UNITW>(P, meta, simacc, ph, big, L, rowidx, sa, vmask, 1, lane, acc);
        prev = ph;
        prev_m = m;
        __syncwarp();
        if (lane == 0) mbar_arrive(sbase + SM_MBAR + 16 + b * 8u);  // this warp is done with buffer b (and its meta record)
    }
    flush_acc();
    if (WANT_SIM) {
        consumer_sync<NC_RES>();
        if (warp == 1) flush_sim(P, sbase + SM_SIMACC + ((p - 1u) & 1u) * 64u, prev, prev_m, lane);
    }
}

// The prefetching consumer (CA_TL_PREFETCH=1, needs the registers of a smaller block): a second register set holds a whole
// super-tile loaded one tile AHEAD.
template <bool WANT_ECC, bool WANT_SIM, bool UNITW>
__device__ __forceinline__ void consume_resident_prefetch(const TileParams& P, const uint32_t sbase, const int tid, const int lane, const int warp) {
    constexpr bool PF = true;
    int m = -1;
    int4 cell = make_int4(-1, -1, -1, -1);
    uint32_t vmask = 0;
    bool wactive = false;
    uint32_t rowidx = 0, sa = 0;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    uint4 L[4];           // labels of the current 8-partner tile for this thread's quad
    uint4 A[4], B[4];     // PF: a whole super-tile (A = partners [16 s, 16 s + 8), B = [16 s + 8, 16 s + 16)), loaded one tile ahead;
                          // otherwise only B is used (the low tile is loaded straight into L)
    int n_st = -1;        // PF: the super-tile whose labels A / B hold for the CURRENT cells (-1 = none)
    bool n_next = false;  // PF: A / B were loaded with the NEXT work unit's cells (they become current at its first phase)
    int4 cellN = make_int4(-1, -1, -1, -1);  // PF: the next work unit's cell ids and quad info, fetched during this one's last tiles
    uint2 qiN = make_uint2(0u, 0u);
    bool haveN = false;
    uint32_t wurec = sbase + SM_WUREC;  // PF: the current unit's record (its last partner and the next unit's coordinates are re-read from it)
    int cur_j0 = -1;
    Phase prev;
    prev.j0 = prev.g0 = prev.cnt = 0;
    int prev_m = -1;

    auto flush_acc = [&]() {
        if (WANT_ECC && m >= 0 && !CA_DBG(P, 0)) {
            const double wscale = __ldg(P.w + m) * P.inv_norm;
            if (cell.x >= 0) ecc_add(P, cell.x, acc[0] * wscale);
            if (cell.y >= 0) ecc_add(P, cell.y, acc[1] * wscale);
            if (cell.z >= 0) ecc_add(P, cell.z, acc[2] * wscale);
            if (cell.w >= 0) ecc_add(P, cell.w, acc[3] * wscale);
        }
    };

    uint32_t p = 0;
    for (;; p++) {
        const uint32_t b = p & 1u;
        const uint32_t meta = sbase + SM_META + b * META_BYTES, simacc = sbase + SM_SIMACC + b * 64u;
        mbar_wait(sbase + SM_MBAR + b * 8u, (p >> 1) & 1u);  // buffer b initialised, meta of phase p published
        const PhaseHdr h = read_hdr(sbase, meta);
        const Phase ph = h.ph;
        if (ph.cnt == 0) break;
        if (h.first) {  // a new work unit: retire the old one, pick up the new item
            flush_acc();
            const uint2 w0 = lds_u64(h.wurec + WU_M), w1 = lds_u64(h.wurec + WU_NPOS), w2 = lds_u64(h.wurec + WU_SA);
            m = (int)w0.x;
            const int npos = (int)w1.x;
            const int64_t off = (int64_t)w2.y * P.perm_stride + (int)w0.y;
            uint2 qi = make_uint2(0u, 0u);
            if (PF && haveN) {
                cell = cellN;
                qi = qiN;
            } else {
                cell = make_int4(-1, -1, -1, -1);
                if (tid * 4 < npos) {
                    cell = __ldg(reinterpret_cast<const int4*>(P.perm + off + tid * 4));
                    qi = __ldg(reinterpret_cast<const uint2*>(P.qinfo) + (off >> 2) + tid);
                }
            }
            if (PF) {
                wurec = h.wurec;
                if (!n_next) n_st = -1;  // what A / B hold belongs to the previous unit's cells
                n_next = false;
                haveN = false;
            }
            wactive = (tid & ~31) * 4 < npos;  // warp-uniform: the warp owns at least one position of the item
            vmask = (cell.x >= 0 ? 1u : 0u) | (cell.y >= 0 ? 2u : 0u) | (cell.z >= 0 ? 4u : 0u) | (cell.w >= 0 ? 8u : 0u);
            rowidx = qi.x;
            sa = qi.y;
            acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
            cur_j0 = -1;
        }
        // ---- labels of the tile ----
        if (ph.j0 != cur_j0) {
            const int st = ph.j0 >> 4;
            if (PF) {
                if (n_st != st) {  // not loaded ahead (the block's first unit, or a unit of a single tile before this one)
                    load_quad_super(P.lab16 + lab16_index(P.n, 0, st << 4), (uint32_t)P.n, cell, A, B);
                    n_st = st;
                }
                const bool hi = (ph.j0 & 8) != 0;
#pragma unroll
                for (int c = 0; c < 4; c++) L[c] = hi ? B[c] : A[c];
                const uint2 w3 = lds_u64(wurec + WU_JHI);  // {the unit's last partner + 1, the next unit's first super-tile}
                const int ti = ph.j0 >> 3, tl = ((int)w3.x - 1) >> 3;  // this tile, the unit's last tile
                const int nxt_st = (int)w3.y;
                const bool hadN = haveN;
                const int nxt_npos = (int)lds_u32(wurec + WU_NXT_NPOS);
                if (ti + 1 >= tl && !haveN && nxt_npos > 0) {  // the unit's last two tiles: fetch the next unit's cell ids and quad info
                    const uint2 w4 = lds_u64(wurec + WU_NXT_SLOT);
                    const int64_t nxt_off = (int64_t)w4.x * P.perm_stride + (int)w4.y;
                    cellN = make_int4(-1, -1, -1, -1);
                    qiN = make_uint2(0u, 0u);
                    if (tid * 4 < nxt_npos) {
                        cellN = __ldg(reinterpret_cast<const int4*>(P.perm + nxt_off + tid * 4));
                        qiN = __ldg(reinterpret_cast<const uint2*>(P.qinfo) + (nxt_off >> 2) + tid);
                    }
                    haveN = true;
                }
                if (hi || ti >= tl) {  // A / B are free from here on: load one tile ahead
                    if (ti < tl) {     // the unit's next super-tile, same cells
                        load_quad_super(P.lab16 + lab16_index(P.n, 0, (st + 1) << 4), (uint32_t)P.n, cell, A, B);
                        n_st = st + 1;
                    } else if (hadN) {  // the next unit's first super-tile, with the cell ids fetched a tile ago
                        load_quad_super(P.lab16 + lab16_index(P.n, 0, nxt_st << 4), (uint32_t)P.n, cellN, A, B);
                        n_st = nxt_st;
                        n_next = true;
                    }
                }
            } else if (wactive) {
                if (cur_j0 < 0 || ((ph.j0 ^ cur_j0) & ~(SUPER - 1)) != 0) {  // the WU's first phase or its second super-tile
                    if (!CA_DBG(P, 2)) {
                        load_quad_super(P.lab16 + lab16_index(P.n, 0, st << 4), (uint32_t)P.n, cell, L, B);
                    } else {  // tuning builds: no label loads, every cell carries label 0
#pragma unroll
                        for (int c = 0; c < 4; c++) L[c] = B[c] = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                if ((ph.j0 & 8) != 0) {
#pragma unroll
                    for (int c = 0; c < 4; c++) L[c] = B[c];
                }
            }
            cur_j0 = ph.j0;
        }
        // ---- histogram ----
        if (wactive && !CA_DBG(P, 3)) hist_cells<4>(P, meta, ph, L, rowidx, 1);
        if (!CA_DBG(P, 5)) consumer_sync<NC_RES>();  // the tables of phase p are complete; every consumer has left phase p-1
        if (WANT_SIM && warp == 1) flush_sim(P, sbase + SM_SIMACC + (b ^ 1u) * 64u, prev, prev_m, lane);
        // ---- gather ----
        if (wactive && !CA_DBG(P, 4)) gather_cells<4, false, WANT_SIM, UNITW>(P, meta, simacc, ph, h.big, L, rowidx, sa, vmask, 1, lane, acc);
        prev = ph;
        prev_m = m;
        __syncwarp();
        if (lane == 0) mbar_arrive(sbase + SM_MBAR + 16 + b * 8u);  // this warp is done with buffer b (and its meta record)
    }
    flush_acc();
    if (WANT_SIM) {
        consumer_sync<NC_RES>();
        if (warp == 1) flush_sim(P, sbase + SM_SIMACC + ((p - 1u) & 1u) * 64u, prev, prev_m, lane);
    }
}

// ---- streaming items: one cluster of more than TL_CAP cells, one table row ----------------------------------------
// The cells are taken in chunks of SCHUNK = 4 * NC_STR positions, a quad per thread and chunk.  An item of at most two
// chunks (most of them) keeps both chunks' cell ids for the whole work unit and their labels between histogram and gather;
// a longer one streams the cells from global memory in both passes.  A warp whose quad of the second chunk lies beyond the
// item skips it (warp-uniform), so the work is counted in warps x chunks, not in whole chunks.  Labels are loaded per
// 8-partner tile (16 bytes per cell); the sums go to ecc per phase.
template <bool WANT_ECC, bool WANT_SIM, bool UNITW, int MODE>
__device__ __forceinline__ void consume_streaming(const TileParams& P, const uint32_t sbase, const int tid, const int lane, const int warp) {
    constexpr int SCHUNK = 4 * NC_STR;
    constexpr bool RT = MODE == MODE_RT, SCAN = MODE == MODE_SCAN;
    const int half = ((P.smem_bytes - SM_DYN) / 2) & ~127;  // RT: a buffer = count tables (a third) + ratio tables
    const uint32_t third = (uint32_t)rt_count_bytes(half);
    int m = -1, npos = 0;
    bool wide = false;
    const int32_t* perm = P.perm;
    uint32_t sa = 0;
    int4 cell = make_int4(-1, -1, -1, -1), cell2 = cell;
    uint4 L[4], Lh[4];
    double wscale = 0.0;
    Phase prev;
    prev.j0 = prev.g0 = prev.cnt = 0;
    int prev_m = -1;
    auto vmask_of = [](int4 c) { return (c.x >= 0 ? 1u : 0u) | (c.y >= 0 ? 2u : 0u) | (c.z >= 0 ? 4u : 0u) | (c.w >= 0 ? 8u : 0u); };
    auto flush4 = [&](int4 cl, const double (&a4)[4]) {
        if (WANT_ECC) {
            if (cl.x >= 0) ecc_add(P, cl.x, a4[0] * wscale);
            if (cl.y >= 0) ecc_add(P, cl.y, a4[1] * wscale);
            if (cl.z >= 0) ecc_add(P, cl.z, a4[2] * wscale);
            if (cl.w >= 0) ecc_add(P, cl.w, a4[3] * wscale);
        }
    };

    uint32_t p = 0;
    for (;; p++) {
        const uint32_t b = p & 1u;
        const uint32_t meta = sbase + SM_META + b * META_BYTES, simacc = sbase + SM_SIMACC + b * 64u;
        mbar_wait(sbase + SM_MBAR + b * 8u, (p >> 1) & 1u);
        const PhaseHdr h = read_hdr(sbase, meta);
        const Phase ph = h.ph;
        if (ph.cnt == 0) break;
        if (h.first) {
            const uint2 w0 = lds_u64(h.wurec + WU_M), w1 = lds_u64(h.wurec + WU_NPOS), w2 = lds_u64(h.wurec + WU_SA);
            m = (int)w0.x;
            npos = (int)w1.x;
            wide = (w1.y & NROWS_WIDE) != 0;
            sa = w2.x;
            perm = P.perm + (int64_t)w2.y * P.perm_stride + (int)w0.y;
            wscale = __ldg(P.w + m) * P.inv_norm;
            if (npos <= 2 * SCHUNK) {
                cell = __ldg(reinterpret_cast<const int4*>(perm + tid * 4));
                cell2 = tid * 4 + SCHUNK < npos ? __ldg(reinterpret_cast<const int4*>(perm + tid * 4 + SCHUNK)) : make_int4(-1, -1, -1, -1);
            }
        }
        const uint16_t* labj = P.lab16 + lab16_index(P.n, 0, ph.j0);
        const bool second = (tid & ~31) * 4 + SCHUNK < npos;  // warp-uniform: the warp owns positions of the second chunk
        // ---- histogram ----
        if (npos <= 2 * SCHUNK) {
            load_quad(labj, (uint32_t)P.n, cell, L);
            if (second) load_quad(labj, (uint32_t)P.n, cell2, Lh);
            hist_cells<4>(P, meta, ph, L, 0u, 1);
            if (second) hist_cells<4>(P, meta, ph, Lh, 0u, 1);
        } else {
            for (int cb = tid * 4; cb < npos; cb += 2 * SCHUNK) {  // two chunks per trip: their loads are in flight together
                const int4 none = make_int4(-1, -1, -1, -1);
                const int4 ca = __ldg(reinterpret_cast<const int4*>(perm + cb));
                const int4 cb4 = cb + SCHUNK < npos ? __ldg(reinterpret_cast<const int4*>(perm + cb + SCHUNK)) : none;
                load_quad(labj, (uint32_t)P.n, ca, L);
                load_quad(labj, (uint32_t)P.n, cb4, Lh);
                hist_cells<4>(P, meta, ph, L, 0u, 1);
                if (cb + SCHUNK < npos) hist_cells<4>(P, meta, ph, Lh, 0u, 1);
            }
        }
        consumer_sync<NC_STR>();
        if (WANT_SIM && warp == 1) flush_sim(P, sbase + SM_SIMACC + (b ^ 1u) * 64u, prev, prev_m, lane);
        // ---- gather ----
        auto gather_item = [&](auto simcell) {  // simcell: the matrix terms are formed per (cell, partner) in the gather
            constexpr bool SC = decltype(simcell)::value;
            if (npos <= 2 * SCHUNK) {
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    if (hh && !second) break;  // warp-uniform
                    const int4 cl = hh ? cell2 : cell;
                    double a4[4] = {0.0, 0.0, 0.0, 0.0};
                    gather_cells<4, false, SC, UNITW>(P, meta, simacc, ph, h.big, hh ? Lh : L, 0u, sa, vmask_of(cl), 1, lane, a4);
                    flush4(cl, a4);
                }
            } else {
                for (int cb0 = 0; cb0 < npos; cb0 += SCHUNK) {  // block-uniform trip count: gather_cells shuffles when SC
                    const int cb = cb0 + tid * 4;
                    int4 cl = make_int4(-1, -1, -1, -1);
                    if (cb < npos) cl = __ldg(reinterpret_cast<const int4*>(perm + cb));
                    load_quad(labj, (uint32_t)P.n, cl, L);
                    double a4[4] = {0.0, 0.0, 0.0, 0.0};
                    if (wide)
                        gather_cells<4, true, SC, UNITW>(P, meta, simacc, ph, h.big, L, 0u, sa, vmask_of(cl), 1, lane, a4);
                    else
                        gather_cells<4, false, SC, UNITW>(P, meta, simacc, ph, h.big, L, 0u, sa, vmask_of(cl), 1, lane, a4);
                    flush4(cl, a4);
                }
            }
        };
        if (RT) {
            const uint32_t cbase = sbase + SM_DYN + b * (uint32_t)half, rbase = cbase + third;
            if (wide)
                convert_tables<true, WANT_SIM>(P, meta, simacc, ph, 1, sa, cbase, rbase, warp - 1, NC_STR / 32, lane);
            else
                convert_tables<false, WANT_SIM>(P, meta, simacc, ph, 1, sa, cbase, rbase, warp - 1, NC_STR / 32, lane);
            consumer_sync<NC_STR>();
            if (npos <= 2 * SCHUNK) {
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    if (hh && !second) break;  // warp-uniform
                    const int4 cl = hh ? cell2 : cell;
                    double a4[4] = {0.0, 0.0, 0.0, 0.0};
                    gather_ratio<UNITW>(P, meta, ph, hh ? Lh : L, 0u, cbase, rbase, a4);
                    flush4(cl, a4);
                }
            } else {
                for (int cb = tid * 4; cb < npos; cb += SCHUNK) {
                    const int4 cl = __ldg(reinterpret_cast<const int4*>(perm + cb));
                    load_quad(labj, (uint32_t)P.n, cl, L);
                    double a4[4] = {0.0, 0.0, 0.0, 0.0};
                    gather_ratio<UNITW>(P, meta, ph, L, 0u, cbase, rbase, a4);
                    flush4(cl, a4);
                }
            }
        } else if (WANT_SIM && SCAN) {  // the matrix from the tables (one row per partner here)
            if (wide)
                scan_sim<true>(P, meta, simacc, ph, 1, sa, warp - 1, NC_STR / 32, lane);
            else
                scan_sim<false>(P, meta, simacc, ph, 1, sa, warp - 1, NC_STR / 32, lane);
            if (WANT_ECC) gather_item(Slot<0>{});
        } else {
            gather_item(Slot<WANT_SIM ? 1 : 0>{});
        }
        prev = ph;
        prev_m = m;
        __syncwarp();
        if (lane == 0) mbar_arrive(sbase + SM_MBAR + 16 + b * 8u);
    }
    if (WANT_SIM) {
        consumer_sync<NC_STR>();
        if (warp == 1) flush_sim(P, sbase + SM_SIMACC + ((p - 1u) & 1u) * 64u, prev, prev_m, lane);
    }
}

template <bool WANT_ECC, bool WANT_SIM, bool UNITW, bool STREAMING, int MODE>
__global__ void __launch_bounds__((STREAMING ? NC_STR : NC_RES) + 32, 1) wave_kernel(const TileParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NCT = STREAMING ? NC_STR : NC_RES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const int half = ((P.smem_bytes - SM_DYN) / 2) & ~127;
    if (threadIdx.x == 0) {
        mbar_init(sbase + SM_MBAR, 1);                 // full[0], full[1]: the producer's arrive + the fill's bytes
        mbar_init(sbase + SM_MBAR + 8, 1);
        mbar_init(sbase + SM_MBAR + 16, NCT / 32);     // empty[0], empty[1]: one arrive per consumer warp
        mbar_init(sbase + SM_MBAR + 24, NCT / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 16) sts_b64(sbase + SM_SIMACC + threadIdx.x * 8u, 0ull);
    __syncthreads();
    if (warp == 0) {
        producer_loop<UNITW>(P, sbase, half, MODE == MODE_RT ? rt_count_bytes(half) : half, lane);
        return;
    }
    const int tid = (int)threadIdx.x - 32;
    if (STREAMING)
        consume_streaming<WANT_ECC, WANT_SIM, UNITW, MODE>(P, sbase, tid, lane, warp);
    else if (CA_TL_PREFETCH && MODE == MODE_PLAIN)
        consume_resident_prefetch<WANT_ECC, WANT_SIM, UNITW>(P, sbase, tid, lane, warp);
    else
        consume_resident<WANT_ECC, WANT_SIM, UNITW, MODE>(P, sbase, tid, lane, warp);
}

// Shared memory one block asks for and the number of table rows a resident block may own so that at least one
// partner with kmax clusters fits into one of the two buffers (R rows of kmax + 1 32-bit entries, rounded to 4).
int tile_smem_budget(int32_t kmax, int32_t* max_rows_out, int32_t* half_out, bool rt) {
    Ctx& c = ctx();
    const size_t optin = c.smem_optin ? c.smem_optin : 227 * 1024;
    const int budget = (int)optin;
    // Leave 32 KB of the SM's 228 KB to the L1 cache (the carve-out steps are ... 132, 164, 196, 228 KB): register
    // spills and the producer's small loads otherwise go to L2 at ~300 cycles each.  Measured on C5 (round 1): 163 KB
    // 210 ms, 179 KB 210 ms, 195 KB 206 ms, 211 KB 237 ms, 227 KB 234 ms.  CA_SMEM_KB overrides it (tuning).
    int budget_kb = 195;
    if (const char* e = getenv("CA_SMEM_KB")) budget_kb = atoi(e);
    const int budget_eff = budget_kb > 0 ? std::min(budget, budget_kb * 1024) : budget;
    const long long half = (((long long)budget_eff - SM_DYN) / 2) & ~127LL;
    const long long rowb = (((long long)(kmax > 0 ? kmax : 1) + 1 + 3) & ~3LL) * 4;  // + the pad column
    long long rows = (rt ? rt_count_bytes((int)half) : half) / rowb;  // ratio-table launches: the count tables get a third of a buffer
    if (rows > TL_MAX_ROWS) rows = TL_MAX_ROWS;
    if (max_rows_out) *max_rows_out = (int32_t)(rows < 0 ? 0 : rows);
    if (half_out) *half_out = (int32_t)half;
    return budget_eff;
}

// per-quad row index and cluster size of every owned partition's cluster-sorted order (streamed by the wave kernel
// instead of three dependent random loads per work unit) -- only behind CA_SORT_LEGACY=1; the chunk sort writes them itself
__global__ void __launch_bounds__(256) qinfo_kernel(const TileParams P, const int32_t* __restrict__ parts, int32_t nparts, int64_t nquads) {
    const int sl = blockIdx.y;
    if (sl >= nparts) return;
    const int m = parts[sl];
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (int64_t)gridDim.x * blockDim.x) {
        const int cellx = __ldg(P.perm + (int64_t)sl * P.perm_stride + q * 4);
        uint2 v = make_uint2(0u, 0u);
        if (cellx >= 0) {
            const int a = (int)__ldg(P.lab16 + lab16_index(P.n, cellx, m));
            v.x = (uint32_t)__ldg(P.crow + (int64_t)sl * P.kp + a);
            v.y = (uint32_t)__ldg(P.sizes + (int64_t)m * P.kp + a);
        }
        reinterpret_cast<uint2*>(P.qinfo)[(int64_t)sl * (P.perm_stride >> 2) + q] = v;
    }
}
int launch_qinfo(const TileParams& p, const int32_t* d_parts, int32_t nparts, cudaStream_t s) {
    if (nparts <= 0) return CA_OK;
    const int64_t nquads = p.perm_stride >> 2;
    unsigned gx = (unsigned)std::min<int64_t>((nquads + 255) / 256, 1024);
    qinfo_kernel<<<dim3(gx, (unsigned)nparts), 256, 0, s>>>(p, d_parts, nparts, nquads);
    CA_CUDA(cudaGetLastError());
    ca::ctx().launches++;
    return CA_OK;
}

template <bool E, bool S, bool U, bool STREAMING, int MODE>
static int launch_wave_k(const TileParams& p, cudaStream_t s) {
    if (p.nwu <= 0) return CA_OK;
    auto kern = wave_kernel<E, S, U, STREAMING, MODE>;
    CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes));
    const int grid = std::min<int>(ca::ctx().sm_count, p.nwu);
    kern<<<(unsigned)grid, (STREAMING ? NC_STR : NC_RES) + 32, (size_t)p.smem_bytes, s>>>(p);
    CA_CUDA(cudaGetLastError());
    ca::ctx().launches++;
    return CA_OK;
}

// p.streaming selects the kind of work units in p.items: the streaming kernel first (long items), then the resident one
template <bool E, bool S, bool U>
static int launch_wave_t(const TileParams& p, cudaStream_t s) {
    // ratio tables need ecc, the matrix from the tables needs the matrix: the other combinations are not instantiated
    constexpr int RTM = E ? MODE_RT : MODE_PLAIN, SCM = S ? MODE_SCAN : MODE_PLAIN;
    if (E && p.rt) return p.streaming ? launch_wave_k<E, S, U, true, RTM>(p, s) : launch_wave_k<E, S, U, false, RTM>(p, s);
    if (S && p.sim_scan) return p.streaming ? launch_wave_k<E, S, U, true, SCM>(p, s) : launch_wave_k<E, S, U, false, SCM>(p, s);
    return p.streaming ? launch_wave_k<E, S, U, true, MODE_PLAIN>(p, s) : launch_wave_k<E, S, U, false, MODE_PLAIN>(p, s);
}

int launch_tile(const TileParams& p, cudaStream_t s) {
    if (p.nwu <= 0) return CA_OK;
    const bool e = p.ecc != nullptr, sm = p.sim != nullptr;
    if (p.unit_w) {
        if (e && sm) return launch_wave_t<true, true, true>(p, s);
        if (e) return launch_wave_t<true, false, true>(p, s);
        if (sm) return launch_wave_t<false, true, true>(p, s);
    } else {
        if (e && sm) return launch_wave_t<true, true, false>(p, s);
        if (e) return launch_wave_t<true, false, false>(p, s);
        if (sm) return launch_wave_t<false, true, false>(p, s);
    }
    return CA_OK;
}

double sim_fix_scale(int64_t n) {  // 2^F, F = min(40, 61 - ceil(log2(n + 1))): n terms of at most 1 stay below 2^62, and v + 1.5 * 2^(52 - F)
    int bits = 0;                  // keeps its binade for every v in [0, 1]
    while (((int64_t)1 << bits) < n + 1) bits++;
    return ldexp(1.0, std::min(40, 61 - bits));
}
double sim_fix_magic(int64_t n) { return 1.5 * 4503599627370496.0 / sim_fix_scale(n); }  // 1.5 * 2^(52 - F)

}  // namespace ca
