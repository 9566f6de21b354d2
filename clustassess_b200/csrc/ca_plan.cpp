// ca_plan.cpp -- host-side item planner (pure C++, no CUDA).
//
// For one partition i with cluster sizes S[0..K) it decides where every cluster sits in the
// cluster-sorted cell order perm_i and which clusters form one ITEM of the wave kernel (ca_tile.cu), i.e. are
// processed together by one thread block and share its shared-memory tables:
//   * every cluster's segment starts at a multiple of `align` (= cells per thread), so a thread never
//     straddles two clusters; the tail of a segment is padding (perm = -1);
//   * a cluster whose padded size exceeds `cap` becomes its own "streaming" item (one table row,
//     cells processed in several sub-chunks; 32-bit counters if it has more than 65535 cells);
//   * all other clusters are packed best-fit-decreasing into "resident" items of at most `cap`
//     positions and at most `max_rows` table rows (their labels live in registers for a whole work unit).
// The reference has no counterpart: it re-scans both label vectors for every pair
// (src/calculate_ecs.cpp:23-26); this planning is what lets one block own complete table rows.
// Note: the wave kernel's fill traffic and phase count grow with the number of rows, whatever the packing.
#include <algorithm>
#include <numeric>

#include "ca_internal.h"

namespace ca {

int plan_blocks(const int32_t* sizes, int32_t k, int32_t align, int32_t cap, int32_t max_rows, PlanOut& out) {
    if (k < 0 || align <= 0 || cap < align || cap % align != 0 || max_rows < 1) {
        set_error("plan_blocks: bad arguments (k=%d align=%d cap=%d max_rows=%d)", k, align, cap, max_rows);
        return CA_E_ARG;
    }
    out.cstart.assign(k, 0);
    out.crow.assign(k, 0);
    out.pos0.clear();
    out.npos.clear();
    out.nrows.clear();
    out.wide.clear();
    out.sa.clear();
    out.total_pos = 0;

    // Everything below is flat arrays reused across calls (one planner per host thread): the planner runs once per
    // owned partition and sits between two GPU stages, so its allocations were what it spent its time on.
    struct Scratch {
        std::vector<uint64_t> keys, tmp;               // cluster indices, sorted by decreasing size
        std::vector<int32_t> head, next, bin_of, bin_rows, bin_start, member, big;
        std::vector<int32_t> bin_left;  // units still free in a bin
        std::vector<uint64_t> nonempty;
    };
    static thread_local Scratch S;

    // `align` is a power of two in every call the library makes (cells per thread): shifts instead of divisions.  (Measured on
    // K = 500 .. 2000 clusters: 43 us per partition with a radix sort and 64-bit vectors in the best-fit loop, 31 us now; the divisions were not it.)
    const bool pow2 = (align & (align - 1)) == 0;
    const int sh = pow2 ? __builtin_ctz((unsigned)align) : 0;
    auto units = [&](int64_t x) -> int64_t { return pow2 ? (x >> sh) : x / align; };   // x / align
    auto padded = [&](int32_t s) -> int64_t { return pow2 ? (((int64_t)s + align - 1) >> sh) << sh : (int64_t)((s + align - 1) / align) * align; };

    // Resident clusters by decreasing PADDED size (what the packing sees), ties by increasing index: a counting sort over the
    // at most cap / align distinct padded sizes -- O(K + cap / align), against a comparison or radix sort of the sizes
    // (17 of the planner's 65 us per partition at K = 1300; the planner is the pre-wave critical path of multi-GPU jobs).
    // The few clusters beyond `cap` (streaming items) are sorted by size on their own.
    const int32_t nb = (int32_t)units(cap) + 1;
    S.keys.resize(k);
    S.tmp.assign((size_t)nb + 1, 0);  // tmp[u] = clusters of u units, then the first slot of that class
    S.big.clear();
    for (int32_t c = 0; c < k; c++) {
        if (sizes[c] <= 0) {
            set_error("plan_blocks: cluster %d has size %d (must be > 0)", c, sizes[c]);
            return CA_E_ARG;
        }
        const int64_t p = padded(sizes[c]);
        if (p > cap)
            S.big.push_back(c);
        else
            S.tmp[(size_t)units(p)]++;
    }
    {
        uint64_t at = 0;
        for (int32_t u = nb - 1; u >= 0; u--) {
            const uint64_t n_u = S.tmp[(size_t)u];
            S.tmp[(size_t)u] = at;
            at += n_u;
        }
    }
    const int32_t nres = k - (int32_t)S.big.size();
    for (int32_t c = 0; c < k; c++) {
        const int64_t u = units(padded(sizes[c]));
        if (u < nb) S.keys[S.tmp[(size_t)u]++] = ((uint64_t)u << 32) | (uint32_t)c;  // {units, cluster}
    }
    std::stable_sort(S.big.begin(), S.big.end(), [&](int32_t x, int32_t y) { return sizes[x] > sizes[y]; });

    // open bins by remaining capacity (in units of `align`): per-capacity LIFO lists threaded through `next` + a bitmap of
    // the non-empty lists, so "smallest remaining capacity that still holds p" is a find-next-set-bit.  Everything in units,
    // arrays sized once: this loop is most of the planner's time.
    const int32_t capu = nb - 1, nwords = (nb + 63) / 64;
    S.head.assign(nb, -1);
    S.nonempty.assign(nwords, 0);
    S.next.resize(std::max(nres, 1));
    S.bin_left.resize(std::max(nres, 1));
    S.bin_rows.resize(std::max(nres, 1));
    S.bin_of.resize(std::max(k, 1));
    int32_t* const head = S.head.data();
    int32_t* const next = S.next.data();
    int32_t* const bin_left = S.bin_left.data();
    int32_t* const bin_rows = S.bin_rows.data();
    int32_t* const bin_of = S.bin_of.data();
    uint64_t* const nonempty = S.nonempty.data();
    int32_t nbins = 0;
    for (int32_t idx = 0; idx < nres; idx++) {
        const uint64_t key = S.keys[idx];
        const int32_t c = (int32_t)(uint32_t)key, pu = (int32_t)(key >> 32);
        int32_t u = -1;  // best fit: the first non-empty list >= pu
        {
            int32_t w = pu >> 6;
            uint64_t bits = nonempty[w] & (~0ULL << (pu & 63));
            while (bits == 0 && ++w < nwords) bits = nonempty[w];
            if (bits) u = w * 64 + __builtin_ctzll(bits);
        }
        int32_t b;
        if (u < 0) {
            b = nbins++;
            bin_left[b] = capu;
            bin_rows[b] = 0;
        } else {
            b = head[u];
            head[u] = next[b];
            if (head[u] < 0) nonempty[u >> 6] &= ~(1ULL << (u & 63));
        }
        bin_of[c] = b;
        const int32_t rem = (bin_left[b] -= pu);
        if (++bin_rows[b] < max_rows && rem >= 1) {
            next[b] = head[rem];
            head[rem] = b;
            nonempty[rem >> 6] |= 1ULL << (rem & 63);
        }
    }
    // members of every bin in the order they were added (decreasing size)
    S.bin_start.assign(nbins + 1, 0);
    for (int32_t b = 0; b < nbins; b++) S.bin_start[b + 1] = S.bin_start[b] + bin_rows[b];
    S.member.resize(S.bin_start[nbins]);
    for (int32_t b = 0; b < nbins; b++) bin_rows[b] = 0;  // reused as the fill cursor
    for (int32_t idx = 0; idx < nres; idx++) {
        const int32_t c = (int32_t)(uint32_t)S.keys[idx];
        const int32_t b = bin_of[c];
        S.member[S.bin_start[b] + bin_rows[b]++] = c;
    }

    int64_t pos = 0;
    auto push_item = [&](int64_t p0, int64_t np, int32_t nr, int32_t wide, int32_t sa) {
        out.pos0.push_back((int32_t)p0);
        out.npos.push_back((int32_t)np);
        out.nrows.push_back(nr);
        out.wide.push_back(wide);
        out.sa.push_back(sa);
    };
    for (int32_t c : S.big) {  // streaming items first: they are the long ones
        int64_t p = padded(sizes[c]);
        out.cstart[c] = (int32_t)pos;
        out.crow[c] = 0;
        push_item(pos, p, 1, sizes[c] > 65535 ? 1 : 0, sizes[c]);
        pos += p;
    }
    for (int32_t b = 0; b < nbins; b++) {
        int64_t p0 = pos;
        int32_t r = 0;
        for (int32_t t = S.bin_start[b]; t < S.bin_start[b + 1]; t++) {
            const int32_t c = S.member[t];
            out.cstart[c] = (int32_t)pos;
            out.crow[c] = r++;
            pos += padded(sizes[c]);
        }
        push_item(p0, pos - p0, r, 0, r == 1 ? sizes[S.member[S.bin_start[b]]] : 0);
    }
    if (pos > 0x7fffff00LL) {
        set_error("plan_blocks: %lld padded positions exceed the int32 position range", (long long)pos);
        return CA_E_UNSUPPORTED;
    }
    out.total_pos = pos;
    return CA_OK;
}

}  // namespace ca

extern "C" int ca_plan_blocks(const int32_t* sizes, int32_t k, int32_t align, int32_t cap, int32_t max_rows,
                              int32_t* cstart, int32_t* crow, int32_t* items, int32_t max_items, int64_t* npos_out) {
    if (!sizes || !cstart || !crow || !items || !npos_out) {
        ca::set_error("ca_plan_blocks: NULL argument");
        return CA_E_ARG;
    }
    ca::PlanOut po;
    int rc = ca::plan_blocks(sizes, k, align, cap, max_rows, po);
    if (rc != CA_OK) return rc;
    int32_t ni = (int32_t)po.pos0.size();
    if (ni > max_items) {
        ca::set_error("ca_plan_blocks: %d items exceed max_items=%d", ni, max_items);
        return CA_E_ARG;
    }
    for (int32_t c = 0; c < k; c++) {
        cstart[c] = po.cstart[c];
        crow[c] = po.crow[c];
    }
    for (int32_t t = 0; t < ni; t++) {
        items[4 * t + 0] = po.pos0[t];
        items[4 * t + 1] = po.npos[t];
        items[4 * t + 2] = po.nrows[t];
        items[4 * t + 3] = po.wide[t];
    }
    *npos_out = po.total_pos;
    return ni;
}
