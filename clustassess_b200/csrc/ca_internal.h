// ca_internal.h -- shared declarations of the B200 ECS library (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/clustassess_b200.h"

namespace ca {

// ---- errors -----------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define CA_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ca::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return CA_E_CUDA;                                                                      \
        }                                                                                          \
    } while (0)
#define CA_TRY(call)            \
    do {                        \
        int rc_ = (call);       \
        if (rc_ != CA_OK) return rc_; \
    } while (0)

// ---- context ----------------------------------------------------------------------------------------
struct Ctx {
    bool inited = false;
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;        // max dynamic shared memory per block (opt-in)
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // the stream everything is launched on
    cudaStream_t up_stream = nullptr;  // host -> device copies that run beside the kernels (pipelined host path)
    cudaEvent_t up_ev[8] = {};         // one per upload group
    cudaStream_t peer_stream[4] = {};  // multi-GPU jobs: DMA copies of this device's label slabs into the peers' memory
    cudaEvent_t peer_ev[4] = {};
    cudaEvent_t ev[10] = {};
    double prof[10] = {};
    long launches = 0;  // kernels launched since the counter was last reset
};
Ctx& ctx();          // the context of the device the calling thread works on (thread-local selection, default: the primary device)
int ensure_init();

// Grow-only scratch buffers every device owns one set of (released by ca_shutdown): what used to be function-local statics.
struct DevBuf;
enum ScratchId { SCR_ZEROS, SCR_ECC, SCR_SIM, SCR_OUT, SCR_SIG, SCR_FLAG, SCR_FIRST, SCR_SEGCUR, SCR_PAIR_A, SCR_PAIR_B, SCR_PAIR_TABLE,
                 SCR_PAIR_SA, SCR_PAIR_SB, SCR_PAIR_ERR, SCR_PAIR_ECS, SCR_PAIR_PARTIAL, SCR_PAIR_OUT, SCR_HASH, SCR_WUCTR, SCR_COUNT };
DevBuf& scratch(int which);

// device memory comes from a small caching allocator (ca_api.cu): freed blocks are parked and reused
cudaError_t dev_alloc(void** out, size_t bytes);
void dev_free(void* p, size_t bytes);
void dev_cache_release();

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <typename T> T* as() { return static_cast<T*>(p); }
};

// ---- work items -------------------------------------------------------------------------------------
// A thread owns a quad (TL_CPT) of consecutive positions of a partition's cluster-sorted cell order; a block of
// TL_THREADS threads owns up to TL_CAP positions.
struct Item {       // a set of complete clusters of partition m: positions [pos0, pos0+npos) of its cluster-sorted order
    int32_t m;      // partition whose table rows this item owns
    int32_t pos0;   // first padded position (multiple of TL_CPT) in perm[slot]
    int32_t npos;   // padded positions (multiple of TL_CPT); > TL_CAP => single-cluster streaming item
    int32_t nrows;  // table rows (clusters); bit 30 = more than 65535 cells, 32-bit counters
    int32_t jbeg;   // partner partitions j in [jbeg, jend)
    int32_t jend;
    int32_t sa;     // cluster size when nrows == 1 (else 0)
    int32_t pad_;   // perm slot of partition m
};

// host planner (ca_plan.cpp)
struct PlanOut {
    std::vector<int32_t> cstart, crow;      // per cluster
    std::vector<int32_t> pos0, npos, nrows, wide, sa;  // per item
    int64_t total_pos = 0;
};
int plan_blocks(const int32_t* sizes, int32_t k, int32_t align, int32_t cap, int32_t max_rows, PlanOut& out);

// ---- prepared label set -----------------------------------------------------------------------------
struct Labels {
    int64_t n = 0;
    int32_t m = 0;             // partitions held HERE: the whole set, or one device's share of a sharded set
    int32_t* d_raw = nullptr;  // m x N int32 (partition-major)
    bool owns_raw = true;      // false: the caller's device memory (ca_labels_wrap)
    bool ranked = false;    // min/max, ranks, compact sizes are valid
    bool prepared = false;  // ... and the uint16 label matrix (super-tile-major, see lab16_index), templates, flags
    // filled by prepare():
    std::vector<int32_t> h_min, h_max, h_k;   // per local partition
    std::vector<int64_t> h_roff;  // offset of partition p's range in d_cnt / d_rank
    bool identity = true;         // every partition's labels are already dense (rank == label - min)
    int32_t kmax = 0, kp = 0, mpad = 0;   // over the WHOLE set (all shards)
    DevBuf d_min, d_max, d_cnt, d_rank, d_kcount, d_roff, d_sizes;   // per local partition
    DevBuf d_lab16, d_spack, d_kinfo, d_tmpl;                        // over the whole set (global partition ids)
    int64_t tmpl_bytes = 0;  // bytes of one partition's table template in d_tmpl (= one table buffer of the wave kernel)
    std::vector<int32_t> h_sizes;  // [m][kp], local partitions
    // ---- one device's share of a sharded set (one process, several GPUs) ----
    int32_t m_glob = 0;            // partitions of the whole set; 0 = this IS the whole set
    std::vector<int32_t> gid;      // global id of local partition k (ascending)
    std::vector<int32_t> spans;    // the 32-partition spans owned (ascending): local block y of the transposition is span spans[y]
    DevBuf d_span, d_sizes_g, d_kcount_g;  // span ids on the device; cluster sizes [m_glob][kp] and counts [m_glob] of ALL partitions
    int dev_index = 0;             // which of the selected devices owns these buffers
    // Several shares on ONE device (the pipelined host path: later shares are still crossing PCIe while earlier ones compute)
    // use the whole-set arrays of a common owner instead of their own
    const Labels* whole = nullptr;
    int32_t mg() const { return m_glob ? m_glob : m; }
    int32_t gid_of(int32_t k) const { return m_glob ? gid[k] : k; }
    const Labels& W() const { return whole ? *whole : *this; }
    const int32_t* sizes_all() const { return static_cast<const int32_t*>(m_glob ? W().d_sizes_g.p : d_sizes.p); }
    const int32_t* kcount_all() const { return static_cast<const int32_t*>(m_glob ? W().d_kcount_g.p : d_kcount.p); }
    const uint16_t* lab16_all() const { return static_cast<const uint16_t*>(W().d_lab16.p); }
    const uint32_t* tmpl_all() const { return static_cast<const uint32_t*>(W().d_tmpl.p); }
    const int32_t* kinfo_all() const { return static_cast<const int32_t*>(W().d_kinfo.p); }
    int64_t tmpl_bytes_all() const { return W().tmpl_bytes; }
};

int prepare_ranks(Labels& L);   // min/max, dense ranks, cluster sizes: all the dedup pre-pass needs (any cluster count)
int prepare_labels(Labels& L);  // ... + uint16 label matrix, table templates: the wave engine's inputs (CA_E_UNSUPPORTED beyond its limits)
bool engine_supports(int32_t kmax);  // cluster counts the wave engine can take (uint16 labels, one table row per buffer)

// Compact labels live in a uint16 array laid out super-tile-major: the 16 partitions [16 t, 16 t + 16) of cell c
// are the 32 contiguous bytes at element index (t * (n + 1) + c) * 16.  All cells' labels for one super-tile are
// thus one DENSE 32 (n + 1)-byte slab (an L2 line holds 4 cells), which is what lets a whole super-tile stay
// L2-resident while every owner partition sweeps over it; with a cell-major [n][mpad] layout the same sectors sit
// 2 mpad bytes apart, each in its own 128-byte L2 line, and the effective footprint is four times larger than the
// cache.  Cell n is the PAD CELL: its label in partition p is K_p, one past the last real cluster.  Padding
// positions of the cluster-sorted order read it, so their histogram increments land in a spare table column that
// no real cell ever gathers from, and the inner loops need no validity tests.
__host__ __device__ inline int64_t lab16_index(int64_t n, int64_t cell, int32_t p) { return ((int64_t)(p >> 4) * (n + 1) + cell) * 16 + (p & 15); }

// kernels' host launchers -----------------------------------------------------------------------------
// ca_pack.cu
int launch_fill_i32(int32_t* x, int64_t n, int32_t v, cudaStream_t s);
int launch_minmax(const int32_t* raw, int64_t n, int32_t m, int32_t* d_min, int32_t* d_max, cudaStream_t s);
int launch_count(const int32_t* raw, int64_t n, int32_t m, const int32_t* d_min, const int64_t* d_roff,
                 const int32_t* h_range, int32_t max_range, uint32_t* d_cnt, cudaStream_t s);
int launch_rank(const uint32_t* d_cnt, const int64_t* d_roff, const int32_t* d_min, const int32_t* d_max, int32_t m,
                uint32_t* d_rank, int32_t* d_kcount, cudaStream_t s);
int launch_sizes(const uint32_t* d_cnt, const uint32_t* d_rank, const int64_t* d_roff, const int32_t* d_min,
                 const int32_t* d_max, int32_t m, int32_t max_range, int32_t kp, int32_t* d_sizes, cudaStream_t s);
int launch_spack(const int32_t* d_sizes, const int32_t* d_kcount, int32_t m, int32_t kp, uint32_t* d_spack, int32_t* d_kinfo, cudaStream_t s);
int launch_table_template(const uint32_t* d_spack, const int32_t* d_kcount, int32_t m, int32_t kp, int64_t tmpl_bytes, int32_t max_rows,
                          uint32_t* d_tmpl, cudaStream_t s);
constexpr int CA_MAX_DEVICES = 16;
struct LabDst {  // the lab16 arrays a transposition stores into: the device's own first, then its peers' (sharded label sets)
    uint16_t* p[CA_MAX_DEVICES];
    int n;
};
struct SizeDst {  // the whole-set cluster-size and cluster-count arrays of every device (sharded label sets)
    int32_t* sizes[CA_MAX_DEVICES];
    int32_t* kcount[CA_MAX_DEVICES];
    int n;
};
int launch_scatter_sizes(const int32_t* sizes_loc, const int32_t* kcount_loc, const int32_t* d_span_of_block, int32_t m_loc, int32_t kp, const SizeDst& dst,
                         cudaStream_t s);
int launch_relabel_transpose(const int32_t* raw, int64_t n, int32_t m, int32_t nblocks, const int32_t* d_span_of_block, const int32_t* d_min,
                             const uint32_t* d_rank, const int64_t* d_roff, const int32_t* d_kcount, const LabDst& dst, cudaStream_t s);
int launch_sort(const int32_t* raw, int64_t n, const int32_t* d_min, const uint32_t* d_rank, const int64_t* d_roff,
                const int32_t* d_kcount, const int32_t* d_cstart, int32_t kp, const int32_t* d_parts, int32_t nparts, int32_t* d_perm,
                int64_t perm_stride, const uint16_t* d_crow, const int32_t* d_sizes, uint32_t* d_qinfo, bool* qinfo_done, cudaStream_t s);

// ca_tile.cu -- the all-pairs engine: persistent blocks (one per SM) of consumer warps + 1 producer warp.
// Resident items (complete clusters, <= TL_CAP positions): TL_THREADS consumer threads x a quad of cells.
// Streaming items (one cluster with more than TL_CAP cells): TL_STR_THREADS consumer threads, chunks of a quad per thread.
// Each kind has its own launch, block size and register allocation.
constexpr int TL_CPT = 4;
#ifndef CA_TL_THREADS
#define CA_TL_THREADS 864
#endif
#ifndef CA_TL_PREFETCH
#define CA_TL_PREFETCH 0   // resident kernel: the next super-tile's labels (and the next work unit's cell ids) are loaded one tile ahead
#endif
#ifndef CA_TL_STR_THREADS
#define CA_TL_STR_THREADS 608
#endif
constexpr int TL_THREADS = CA_TL_THREADS;          // tuning builds (tools/build_variant.sh) override these
constexpr int TL_STR_THREADS = CA_TL_STR_THREADS;
constexpr int TL_BLOCKS = 1;                       // persistent blocks per SM
constexpr int TL_CAP = TL_CPT * TL_THREADS;        // positions per resident item
constexpr int TL_MAX_ROWS = 1024;
constexpr int TL_SUPER = 16;                       // partners per super-tile: one 32-byte label sector per cell
#ifndef CA_TL_WU_SPAN
#define CA_TL_WU_SPAN 32
#endif
constexpr int TL_WU_SPAN = CA_TL_WU_SPAN;    // partners per work unit: 1 or 2 super-tiles (cell ids, quad info and the ecc flush are per work unit)
static_assert(TL_WU_SPAN == 16 || TL_WU_SPAN == 32, "a work unit spans one or two super-tiles");
// Per-cell sums and per-pair sums are accumulated as 64-bit FIXED-POINT integers: integer addition is associative, so
// the results are bit-identical from run to run (and across any sharding over GPUs) although the adds are atomics issued
// in arbitrary order.  ecc: units of 2^-61 of the normalised sum (<= 1); a work unit's sum is rounded once, error
// <= 2^-62 per add, far below the 1e-9 bound.  sim: units of 2^-F, F = min(40, 61 - ceil(log2(n + 1))), of sum_n ECS <= n; every
// TERM is rounded to that grid before it is added (error <= 2^-(F+1) per cell, so <= 2^-(F+1) in the mean: 4.5e-13 up to 2M cells).
constexpr double ECC_FIX_SCALE = 2305843009213693952.0;  // 2^61
struct TileParams {
    const uint16_t* lab16;    // [mpad/16][n + 1][16] compact labels (cell n = the pad cell), see lab16_index()
    int64_t n;
    const int32_t* perm;      // [m_slots][perm_stride] cluster-sorted cell ids, -1 = padding
    const int32_t* sizes;     // [m][kp] cluster sizes
    const uint32_t* tmpl;     // [m][tmpl_bytes / 4] table templates: partition p's row (min(size, 65535) << 16 per cluster, 1 << 16 in the
                              // pad column) repeated max_rows times at the table's row stride: one bulk copy initialises an R-row table
    int64_t tmpl_bytes;
    const int32_t* kinfo;     // [mpad] K_j | (1 << 30 if partition j has a cluster with >= 65535 cells), 0 beyond m
    const void* zeros;        // >= 256 KB of zero bytes: source of the bulk copy that clears the tables
    const uint16_t* crow;     // [m_slots][kp] row of a cluster inside its block
    const double* w;          // [mpad + 8] weights
    const Item* items;        // sorted by jbeg; nrows carries bit 30 = 32-bit counters, pad_ = perm slot
    const int32_t* wu_prefix; // [nsuper + 1] work units before span s (work unit = item x TL_WU_SPAN partners)
    uint32_t* qinfo;          // [m_slots][perm_stride / 4] x {row in block, cluster size} of every quad of positions
    unsigned long long* ecc;  // [n] fixed-point partial sums (units of 2^-61), accumulated with integer atomics (may be null)
    unsigned long long* sim;  // [m*m] fixed-point sums of ECS over the cells (units of 1 / sim_fix_scale(n)) (may be null)
    double sim_magic;         // 1.5 * 2^(52 - F): (v + sim_magic) - sim_magic rounds v to the grid 2^-F (sim_fix_magic(n))
    double sim_scale;         // 2^F (sim_fix_scale(n))
    int32_t rt;               // ratio tables: one division per table entry, the gather adds finished ratios (few clusters; ca_tile.cu convert_tables)
    int32_t sim_scan;         // the matrix entries are summed over the table entries (few clusters) instead of per (cell, partner), ca_tile.cu scan_sim
    int64_t perm_stride;
    int32_t m, mpad, kp;
    int32_t nsuper, nwu;      // super-tiles, work units
    int32_t* wu_counter;      // zeroed before the launch: work units beyond the first gridDim.x are handed out in order, first come first served (null: round-robin)
    int32_t streaming;        // kind of the items of this launch: 0 resident (<= TL_CAP positions), 1 streaming
    int32_t smem_bytes;       // dynamic shared memory per block
    double inv_norm;
    int32_t unit_w;           // every weight is exactly 1
    uint32_t sel[2];          // PRMT selectors of a label word's low / high half (0x4410, 0x4432), see slot_label()
    int32_t dbg;              // ablation switches of -DCA_ABLATE tuning builds (tools/build_variant.sh), else 0
};
int launch_tile(const TileParams& p, cudaStream_t s);
int launch_qinfo(const TileParams& p, const int32_t* d_parts, int32_t nparts, cudaStream_t s);
int tile_smem_budget(int32_t kmax, int32_t* max_rows_out, int32_t* half_out = nullptr, bool rt = false);

// ca_pair.cu
int pair_contingency(const int32_t* d_a, const int32_t* d_b, int64_t n, int32_t min_a, int32_t min_b, int32_t ka,
                     int32_t kb, uint32_t* d_table, uint32_t* d_sa, uint32_t* d_sb, int32_t* d_err, cudaStream_t s);
int pair_ecs_gather(const int32_t* d_a, const int32_t* d_b, int64_t n, int32_t min_a, int32_t min_b, int32_t ka,
                    const uint32_t* d_table, const uint32_t* d_sa, const uint32_t* d_sb, double* d_ecs,
                    double* d_partial, int32_t* nblocks_out, cudaStream_t s);
int pair_mean_from_table(const uint32_t* d_table, const uint32_t* d_sa, const uint32_t* d_sb, int32_t ka, int32_t kb,
                         int64_t n, double* d_partial, int32_t* nblocks_out, cudaStream_t s);
int reduce_partials(const double* d_partial, int32_t nblocks, double scale, double* d_out, cudaStream_t s);
int launch_finish(double* ecc, int64_t n, double c0, double* sim, int32_t m, cudaStream_t s);  // fixed-point -> double, in place
// one pair with any cluster counts: hashed contingency table in global memory (the pair-loop fallback beyond the engine's limits)
int pair_hashed_accumulate(const int32_t* d_a, const int32_t* d_b, int64_t n, const int32_t* d_min_a, const int32_t* d_min_b,
                           const uint32_t* rk_a, const uint32_t* rk_b, const int32_t* sizes_a, const int32_t* sizes_b, double scale,
                           unsigned long long* ecc, unsigned long long* sim_slot, double sim_magic, cudaStream_t s);
struct PeerSrc {  // the peers' partial-sum buffers, read through peer memory
    const unsigned long long* p[CA_MAX_DEVICES];
    int n;
};
int launch_peer_sum(unsigned long long* dst, const PeerSrc& src, int64_t n, cudaStream_t s);
double sim_fix_scale(int64_t n);  // 2^F, F = min(40, 61 - ceil(log2(n + 1)))
double sim_fix_magic(int64_t n);  // 1.5 * 2^(52 - F)

// ca_dedup.cu
int dedup_signatures(Labels& L, std::vector<uint64_t>& sig_lo, std::vector<uint64_t>& sig_hi, DevBuf& d_first, cudaStream_t s);
int dedup_same_partition(Labels& L, const DevBuf& d_first, int32_t p, int32_t q, bool* same, cudaStream_t s);

}  // namespace ca
