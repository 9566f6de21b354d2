/*
 * clustassess_b200.h -- C ABI of the B200-native element-centric-similarity (ECS) library.
 *
 * This is the drop-in boundary for ClustAssess's flat-disjoint ECS hot path.  The R package reaches its
 * native code through .Call (reference: R/RcppExports.R:4-14 -> src/RcppExports.cpp:14-51, registration
 * table src/RcppExports.cpp:129-145); a ~100-line C shim (r/clustassess_shim.c)
 * forwards those .Call entries to the functions below.  Plain pointers and sizes only: no R, Rcpp or
 * torch types cross this boundary.
 *
 * Conventions
 *   - Every function returns CA_OK (0) or a negative CA_E_* code; ca_last_error() gives the message.
 *     Nothing here throws, longjmps or calls back into the host language (reference error convention:
 *     BEGIN_RCPP/END_RCPP at src/RcppExports.cpp:17,26 -- the shim raises Rf_error after we return).
 *   - Labels are int32 (R INTSXP / factor codes; the shim truncates REALSXP like Rcpp's
 *     input_parameter<IntegerVector>, src/RcppExports.cpp:20-21).  NA_integer_ (INT_MIN) is rejected with
 *     CA_E_NA_LABEL (the reference has undefined behaviour there).
 *   - A "label matrix" is partition-major M x N: partition p occupies labels[p*n .. p*n+n), which is
 *     R's list-of-vectors layout.  The *_cols variants take M separate column pointers instead.
 *   - Host buffers may be pageable (R vectors are): uploads of 8 MB and more from pageable memory are staged
 *     through pinned slots by several host threads (near the PCIe rate); pinned or registered buffers are copied
 *     directly.
 *   - Inputs are never written.  Outputs are caller-allocated; "host" pointers unless the name ends in
 *     _dev.  There is NO CPU fallback: without a CUDA device every compute entry returns CA_E_NO_DEVICE.
 *   - fp64 throughout.  Per-pair ECS (ca_ecs_pair, ca_ecs_pair_mean, ca_contingency) is one IEEE division per cell:
 *     bit-identical to the reference.  The BATCHED entries evaluate count / max as count * (1 / max) with a
 *     Newton-refined reciprocal (relative error < 1e-13) and sum in another order than the reference: set-level
 *     results (ecc, matrix, agreement) are within 1e-9 absolute of it, not bit-identical to it.
 *   - Set-level results are REPRODUCIBLE: the sums are accumulated as 64-bit fixed-point integers, so two runs of the same
 *     call -- on one GPU or sharded over several, from resident labels or from host buffers -- return identical bits.
 *     (Different calls, e.g. with and without the matrix or with and without weights, run different kernel instantiations
 *     and may differ from each other in the last bits, far below the 1e-9 bound.)
 *   - Host buffers: jobs of 512 MB of labels and more on one device overlap their upload with the computation (the partitions
 *     are uploaded in groups from the last one down; owner i only meets partners j > i).  CA_NO_PIPELINE=1 disables that,
 *     CA_PIPE_MIN_MB moves the threshold.
 *   - Cluster counts: the batched engine takes up to 24 799 clusters per partition; beyond that (any count that fits
 *     memory, like the reference) the single-device entries loop over the pairs with a hashed contingency table.
 */
#ifndef CLUSTASSESS_B200_H
#define CLUSTASSESS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CA_OK 0
#define CA_E_NO_DEVICE (-1)   /* no usable CUDA device / ca_init failed                         */
#define CA_E_CUDA (-2)        /* a CUDA runtime call failed (message has the CUDA error string)  */
#define CA_E_ARG (-3)         /* bad argument (NULL pointer, negative size, m < 2 where required) */
#define CA_E_NA_LABEL (-4)    /* a label equals INT_MIN (R's NA_integer_)                        */
#define CA_E_RANGE (-5)       /* label outside [min, min+k) passed to ca_contingency              */
#define CA_E_UNSUPPORTED (-6) /* label range or cluster count beyond the engine's limits          */
#define CA_E_NOMEM (-7)       /* host or device allocation failed                                 */

/* ---- lifetime ------------------------------------------------------------------------------------ */

/* Select the device(s) this process computes on.  devices==NULL or ndev<=0 -> device 0.  The same list again is a
 * no-op; a different list releases everything held on the old devices first (refused with CA_E_ARG while label sets
 * created on them are alive).
 * ndev > 1: ONE process drives all the listed GPUs (reference analogue: the foreach %dopar% workers of R/ECS.R:953, but
 * inside one .Call): the batched entries below -- ca_consistency[_cols], ca_sim_matrix[_cols], ca_*_resident -- shard the
 * label matrix by 32-partition spans, one host thread per device packs its share and copies the compact labels into every
 * device's memory over NVLink, every device runs the pairs of its own partitions, and ONE ncclAllReduce (64-bit integer
 * sums) combines the per-cell sums (M x M sums likewise).  Needs peer access between the devices; libnccl.so.2 is loaded
 * at this call (without it the sums are added by a kernel that reads peer memory).  ca_device_count(): devices selected. */
int ca_init(const int *devices, int ndev);
int ca_device_count(void);
void ca_shutdown(void);
const char *ca_last_error(void);
int ca_version(void);

/* Launch the (primary device's) work on a caller's CUDA stream.
 * ca_set_stream(handle): handle is a cudaStream_t passed as void*; NULL restores the library's own stream.
 * ca_use_stream(handle, which) also reaches the streams a NULL handle stands for:
 *   CA_STREAM_LIBRARY the library's own non-blocking stream (default), CA_STREAM_HANDLE the given handle (NULL = the legacy
 *   default stream), CA_STREAM_LEGACY / CA_STREAM_PER_THREAD the CUDA sentinels.
 * Work is ordered on that stream only: whoever produced device memory handed to ca_labels_wrap on ANOTHER stream must
 * synchronise with it first. */
#define CA_STREAM_LIBRARY 0
#define CA_STREAM_HANDLE 1
#define CA_STREAM_LEGACY 2
#define CA_STREAM_PER_THREAD 3
int ca_set_stream(void *cuda_stream);
int ca_use_stream(void *cuda_stream, int which);

/* ---- legacy per-pair entry points (replace the three Rcpp routines one for one) ------------------- */

/* Host scan for min/max of a label vector (R computes min(mb) itself before calling myContTable,
 * R/ECS.R:989).  */
int ca_label_range(const int32_t *x, int64_t n, int32_t *min_out, int32_t *max_out);

/* myContTable(a, b, min_a, min_b) -- src/calculate_ecs.cpp:7-19, .Call entry src/RcppExports.cpp:16.
 * table_out: ka x kb int32, COLUMN-major like the R matrix it fills: T[i + j*ka] = #{n: a_n-min_a == i,
 * b_n-min_b == j}, with ka = max(a)-min_a+1, kb = max(b)-min_b+1 (labels unused inside the range keep
 * all-zero rows/cols, as in the reference).  sizes_a (ka) / sizes_b (kb) may be NULL; they receive the
 * row / column sums (= cluster sizes, calculate_ecs.cpp:31). */
int ca_contingency(const int32_t *a, const int32_t *b, int64_t n, int32_t min_a, int32_t min_b, int32_t ka,
                   int32_t kb, int32_t *table_out, int32_t *sizes_a, int32_t *sizes_b);

/* disjointECS(mb1, mb2) -- src/calculate_ecs.cpp:22-46, .Call entry src/RcppExports.cpp:30.
 * ecs_out[n] = T[a_n][b_n] / max(|A(a_n)|, |B(b_n)|).  mean_out (may be NULL) receives mean(ecs_out),
 * which is what element_sim_matrix / weighted_element_consistency take of it (R/ECS.R:954,1479). */
int ca_ecs_pair(const int32_t *a, const int32_t *b, int64_t n, double *ecs_out, double *mean_out);

/* disjointECSaverage(mb1, mb2) -- src/calculate_ecs.cpp:49-70, .Call entry src/RcppExports.cpp:42.
 * sum T^2 / (n * max(S1, S2)) over non-empty cells; n*max is formed in fp64 (the reference's int*int
 * overflows for n*max >= 2^31, and yields NaN when both ranges contain an unused label). */
int ca_ecs_pair_mean(const int32_t *a, const int32_t *b, int64_t n, double *mean_out);

/* ---- batched entry points (the all-pairs hot path) ----------------------------------------------- */

/* weighted_element_consistency(list, weights, calculate_sim_matrix) -- R/ECS.R:1446-1500, and through it
 * element_consistency / merge_partitions(ecs_thresh = 1) -- R/ECS.R:1350-1386, 1166-1226.
 *   ecc_out[n] = [ sum_i w_i(w_i-1)/2 + sum_{i<j} w_i w_j ECS_ij[n] ] / [ W(W-1)/2 ],  W = sum w.
 * weights == NULL -> all 1 (R/ECS.R:1466-1468).  sim_out (m x m, symmetric, unit diagonal; may be NULL)
 * receives mean(ECS_ij) (R/ECS.R:1479-1482).  m < 2 -> CA_E_ARG (R stop(), R/ECS.R:1451). */
int ca_consistency(const int32_t *labels, int64_t n, int32_t m, const double *weights, double *ecc_out,
                   double *sim_out);
int ca_consistency_cols(const int32_t *const *cols, int64_t n, int32_t m, const double *weights,
                        double *ecc_out, double *sim_out);

/* element_sim_matrix_flat_disjoint(mb_list) -- R/ECS.R:931-981 (matrix output): sim_out m x m,
 * sim[i][j] = mean(ECS_ij), unit diagonal; m == 1 -> the 1x1 matrix 1 (R/ECS.R:938-940). */
int ca_sim_matrix(const int32_t *labels, int64_t n, int32_t m, double *sim_out);
int ca_sim_matrix_cols(const int32_t *const *cols, int64_t n, int32_t m, double *sim_out);

/* element_agreement_flat_disjoint(reference, list) -- R/ECS.R:1625-1648:
 * agr_out[n] = mean_p ECS(reference, labels[p])[n].   */
int ca_agreement(const int32_t *reference, const int32_t *labels, int64_t n, int32_t m, double *agr_out);
int ca_agreement_cols(const int32_t *reference, const int32_t *const *cols, int64_t n, int32_t m, double *agr_out);

/* merge_identical_partitions(list) -- R/ECS.R:1021-1054 with are_identical_memberships R/ECS.R:984-1016:
 * greedy first-match dedup.  keep_out[u] = index of the u-th kept partition, freq_out[u] = summed
 * frequency (freq_in == NULL -> all 1), *u_out = number kept.  Reproduces the reference's quirk that a
 * partition with an unused label inside its range is never identical to anything unless K == n.
 *  */
int ca_merge_identical(const int32_t *labels, int64_t n, int32_t m, const double *freq_in, int32_t *keep_out,
                       double *freq_out, int32_t *u_out);
int ca_merge_identical_cols(const int32_t *const *cols, int64_t n, int32_t m, const double *freq_in,
                            int32_t *keep_out, double *freq_out, int32_t *u_out);

/* ---- device-resident label sets (pipeline glue; also what bench.py times as "value") --------------- */

typedef struct ca_labels ca_labels_t;

/* ca_labels_create: device M x N int32, uninitialised.  With several devices selected (ca_init) and more than 32 partitions the
 * set is SHARDED over them by 32-partition spans; ca_labels_create_on_primary always makes a single-device set (what the dedup
 * pre-pass, ca_consistency_partial_dev and ca_labels_device_ptr need).  ca_labels_shape returns n and m of a set. */
int ca_labels_create(int64_t n, int32_t m, ca_labels_t **out);
int ca_labels_create_on_primary(int64_t n, int32_t m, ca_labels_t **out);
int ca_labels_shape(ca_labels_t *h, int64_t *n_out, int32_t *m_out);
int ca_labels_is_sharded(ca_labels_t *h);
int ca_labels_wrap(int64_t n, int32_t m, void *device_matrix, ca_labels_t **out); /* caller-owned device M x N int32
                                                                  (e.g. filled by an NCCL all-gather)  */
int ca_labels_upload(ca_labels_t *h, const int32_t *labels);   /* host -> device, whole matrix         */
int ca_labels_upload_col(ca_labels_t *h, int32_t p, const int32_t *col);
int ca_labels_upload_cols(ca_labels_t *h, const int32_t *const *cols); /* all M columns, one staged upload     */
void *ca_labels_device_ptr(ca_labels_t *h);                    /* raw device matrix (int32 M x N); NULL for a sharded set */
int ca_labels_invalidate(ca_labels_t *h);                      /* call after writing through the ptr   */
int ca_labels_destroy(ca_labels_t *h);

/* Pipeline glue: the stability pipeline compares the SAME partitions several
 * times -- per k (R/stability-3-graph-clustering.R:271-300: merge_partitions of the runs that gave k clusters),
 * per resolution (:25-69: weighted_element_consistency of the merged partitions) and per k across resolutions
 * (:303-343) -- and today re-walks the R vectors every time.  With a resident label set the runs of one
 * configuration are uploaded once; every comparison names its partitions by index:
 *   ca_labels_select     : a new device-resident set holding partitions idx[0..k) of src (device-to-device copies)
 *   ca_consistency_resident : weighted_element_consistency (R/ECS.R:1446-1500) of a resident set, results to host;
 *                          idx == NULL -> the whole set, else the partitions idx[0..k) (weights has k entries)
 *   ca_sim_matrix_resident  : element_sim_matrix_flat_disjoint (R/ECS.R:931-981) likewise. */
int ca_labels_select(ca_labels_t *src, const int32_t *idx, int32_t k, ca_labels_t **out);
int ca_consistency_resident(ca_labels_t *h, const int32_t *idx, int32_t k, const double *weights, double *ecc_out,
                            double *sim_out);
int ca_sim_matrix_resident(ca_labels_t *h, const int32_t *idx, int32_t k, double *sim_out);
/* merge_identical_partitions (R/ECS.R:1021-1054) of a resident single-device set: element_consistency / merge_partitions upload
 * their partitions once, dedup them here and hand the kept indices + summed frequencies to ca_consistency_resident. */
int ca_merge_identical_resident(ca_labels_t *h, const double *freq_in, int32_t *keep_out, double *freq_out, int32_t *u_out);

/* One rank's share of the all-pairs job on a device-resident label set.
 *   rank/world : this process handles the partitions i with zigzag(i) == rank and all their pairs (i,j>i);
 *                world == 1 -> everything.
 *   ecc_dev    : device buffer of n 8-byte words     (may be NULL)  += sum over owned i of (w_i/norm) sum_{j>i} w_j ECS_ij
 *   sim_dev    : device buffer of m*m 8-byte words   (may be NULL)  += sum_n ECS_ij[n] at [i*m+j] for owned i, j>i
 * The words are 64-bit FIXED-POINT integers until ca_consistency_finish_dev() has run: zero the buffers before the first
 * partial call, sum them over the ranks AS INTEGERS (one NCCL allreduce of int64 each: the result is then independent of the
 * reduction order), then ca_consistency_finish_dev() converts them to doubles in place, adds the identical-pair constant,
 * and mirrors / normalises the matrix. */
int ca_shard_owner(int32_t i, int32_t world); /* rank that owns partition i and its pairs (i, j > i) */
int ca_consistency_partial_dev(ca_labels_t *h, const double *weights_host, int32_t rank, int32_t world,
                               double *ecc_dev, double *sim_dev);
int ca_consistency_finish_dev(int64_t n, int32_t m, const double *weights_host, double *ecc_dev,
                              double *sim_dev);

/* Stage timings (milliseconds, CUDA events on the library's stream) of the last batched call on this
 * thread: [0] pack (min/max, cluster sizes, compaction)  [1] relabel + transpose to the uint16 label
 * layout  [2] host planning (items)  [3] per-partition counting sort  [4] the wave kernel (tables + ECS)
 * [5] finish  [6] whole call  [7] number of kernel launches of the call  [8] the wave kernel's streaming launch
 * [9] its resident launch.  Returns how many entries were written (<= cap). */
int ca_get_profile(double *ms_out, int cap);
int ca_get_profile_dev(int device_index, double *ms_out, int cap); /* the same for the index-th selected device of a multi-GPU job;
                                                                   there [0] = packing incl. the label exchange (host clock) */

/* ---- host-only test hook --------------------------------------------------------------------------- */

/* The item planner, exposed so that CPU tests can check its invariants without a GPU.
 * sizes[k] > 0 for k < K.  align = cells per thread, cap = cells per resident block, max_rows = table rows
 * that fit in shared memory.  Outputs: cstart[k] (first padded position of cluster k), crow[k] (row of
 * cluster k inside its block), items[4*t .. 4*t+3] = {pos0, npos, nrows, wide}.  Returns the number of
 * items (<= max_items) or a negative error; *npos_out = padded positions in total. */
int ca_plan_blocks(const int32_t *sizes, int32_t k, int32_t align, int32_t cap, int32_t max_rows,
                   int32_t *cstart, int32_t *crow, int32_t *items, int32_t max_items, int64_t *npos_out);

/* How a sharded label set deals its 32-partition spans to the devices: owner_out[s] = index of the device that holds span s and
 * runs the pairs (i, j > i) of its partitions i.  Returns the number of spans (<= cap) or a negative error.  Needs no GPU. */
int ca_plan_spans(int32_t m, int32_t ndev, int32_t *owner_out, int32_t cap);

/* Which way through the engine a batched job takes, from its largest cluster count and its number of cells: *ratio_tables_out = one
 * IEEE division per contingency-table entry and a gather that only adds (few, huge clusters; needs ecc); *sim_from_tables_out = the
 * ECS matrix is summed over the non-empty table entries, T^2 / max(S_a, S_b) -- the reference's formula for a pair's average
 * (src/calculate_ecs.cpp:58-67) -- instead of per (cell, partner).  All routes give the same numbers to 1e-12 (different rounding
 * of the last bits); each is bit-reproducible.  CA_RT=0/1 and CA_SIM_SCAN=0/1 in the environment force them.  Needs no GPU. */
int ca_plan_route(int32_t kmax, int64_t n, int want_ecc, int want_sim, int *ratio_tables_out, int *sim_from_tables_out);

#ifdef __cplusplus
}
#endif
#endif /* CLUSTASSESS_B200_H */
