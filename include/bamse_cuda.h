/* libbamse_cuda.so -- C ABI of the B200-native BAMSE tree-likelihood scorer.
 *
 * The reference (HoseinT/BAMSE) is pure Python and has NO plugin / FFI interface
 * (SURVEY.md section 8b).  The entry points below are what a ctypes binding for its hot
 * path would bind; each cites the reference code it replaces (paths relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every `const T*` without a `d_` prefix is
 *    caller-owned HOST memory (C-contiguous), read-only; `out_*` are caller-owned host
 *    buffers the library fills; `d_*` are DEVICE pointers on the ctx's device.
 *  - return 0 on success, a negative BAMSE_E* code on failure; never throws/aborts across
 *    the ABI; bamse_last_error() returns the message of the last failure on the ctx
 *    (or of the last failed bamse_create when ctx == NULL).
 *  - a ctx is bound to ONE CUDA device and is not re-entrant (one host thread at a
 *    time); distinct ctxs may be used concurrently.  Multi-GPU = one process (rank)
 *    per GPU, each with its own ctx, sharding by tree-index range; the only exchange
 *    is an all-gather of bamse_record lists (done by the host with NCCL) followed by
 *    bamse_merge_topk_dev on every rank.
 *  - there is NO CPU fallback: with no usable CUDA device bamse_create fails with
 *    BAMSE_ENODEV.
 *  - a "tree index" t in [0, K^(K-1)) names a labeled rooted tree on K nodes:
 *    root = t % K, and t / K holds the K-2 Pruefer digits in base K, least significant
 *    first (oracle/bamse_oracle.py: decode_tree is the executable definition).
 */
#ifndef BAMSE_CUDA_H
#define BAMSE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAMSE_ABI_VERSION 1
#define BAMSE_GRID 512      /* clustering.py:20,26  hard-coded grid length              */
#define BAMSE_KMAX 12       /* bamse.py:32          default of -nc; 12^11 fits uint64    */
#define BAMSE_MMAX 64       /* samples per problem                                       */
#define BAMSE_TOPK_MAX 64
#define BAMSE_TOPK_INTERNAL 256 /* bamse_score_topk may preselect up to this many per parameter set  */

#define BAMSE_OK 0
#define BAMSE_EINVAL (-1)   /* bad argument                                              */
#define BAMSE_ENODEV (-2)   /* no CUDA device / not sm_100                               */
#define BAMSE_ECUDA (-3)    /* CUDA runtime error (message has the cudaError string)     */
#define BAMSE_ESTATE (-4)   /* call order (e.g. scoring before a problem is uploaded)    */
#define BAMSE_ENOMEM (-5)   /* device or host allocation failed                         */
#define BAMSE_EINTERNAL (-6) /* unexpected internal error (a C++ exception was caught)    */

#define BAMSE_FP32 0        /* production: fp32 arithmetic, fp64 accumulation over samples */
#define BAMSE_FP64 1        /* check mode: everything in fp64, reference operation order  */
#define BAMSE_FP32_LOGDOMAIN 2 /* the check-mode algorithm in fp32 (no structural assumptions) */

typedef struct bamse_ctx bamse_ctx;

/* One ranked candidate.  32 bytes, identical on host and device.
 * total = score + clust_const[param][clust]            (clustering.py:183 `totalscore`)
 * score = get_tree_score + log(scre)                   (clustering.py:179,182)        */
typedef struct bamse_record {
    double total;
    double score;
    int32_t clust;
    int32_t param;
    uint64_t tree;
} bamse_record;

/* ---------------------------------------------------------------- life cycle */
int bamse_abi_version(void);
int bamse_create(bamse_ctx** out, int device_ordinal);
void bamse_destroy(bamse_ctx* ctx);
const char* bamse_last_error(const bamse_ctx* ctx);
/* Launch all work of this ctx on `cuda_stream` (a cudaStream_t; NULL = the ctx's own
 * stream).  Lets a host that owns streams (e.g. torch) time the kernels with its events. */
int bamse_set_stream(bamse_ctx* ctx, void* cuda_stream);
/* Multi-GPU sharding: with (rank, world) set, every enumeration of this ctx scores only the
 * tree indices begin + rank, begin + rank + world, ... of each requested range (interleaved,
 * so that every rank sees the same mix of tree shapes).  Default (0, 1) = everything.
 * bamse_score_range is not affected. */
int bamse_set_shard(bamse_ctx* ctx, int rank, int world);
/* The sharding plan of the resident problem: out_owner [P][C] = the rank that scores the whole (p, c), or -1 when
 * its tree indices are split over all ranks (also when there is no plan). */
int bamse_get_shard_plan(bamse_ctx* ctx, int32_t* out_owner);
/* Which instantiation of the fp32 enumeration kernel the resident problem uses: 0 = exponential
 * (MUFU) path only, 1 = with the linear-domain "framed" path, -1 = not decided yet.  Decided by a
 * ~1 ms timing probe on the first enumeration of a problem (BAMSE_FRAMED=0/1 overrides). */
int bamse_fast_variant(const bamse_ctx* ctx);
/* Options.  BAMSE_OPT_EARLY_EXIT (default 0): in fp32 top-k enumerations, stop scoring a tree as soon
 * as its partial sum over samples is below what it needs (the threshold of its clustering / the k-th
 * best so far).  Exact: every per-sample score is <= 0, so a tree stopped early cannot qualify, and
 * the returned top-k is identical (it is re-scored in fp64 either way).  This mirrors what the
 * reference's branch-and-bound does with partial forests (clustering.py:136-138), but with a bound
 * that never discards a qualifying tree.  Off by default: bench.py's headline scores every sample of
 * every tree.  bamse_samples_evaluated returns the (tree, sample) evaluations since the last reset. */
#define BAMSE_OPT_EARLY_EXIT 1
/* BAMSE_OPT_SHARD_PLAN (default 1): with a shard (rank, world > 1) set BEFORE bamse_upload_problem, whole
 * (parameter set, clustering) pairs are assigned to ranks by estimated cost (tables are built only where they are
 * needed) and only the most expensive pairs are split over all ranks (by pack node sets on the pair path, by interleaved tree
 * index otherwise); 0 = split
 * everything.  Every rank computes the same plan.  The plan belongs to the resident problem: changing the shard
 * afterwards requires a new upload (scoring calls fail with BAMSE_ESTATE until then).
 * BAMSE_OPT_STATS (default 0): the table-building kernels also count their executed work (bamse_kernel_stats). */
#define BAMSE_OPT_SHARD_PLAN 2
#define BAMSE_OPT_STATS 3
int bamse_set_option(bamse_ctx* ctx, int option, int value);
int64_t bamse_samples_evaluated(bamse_ctx* ctx, int reset);
/* bamse_score_topk selects a short list with the enumeration kernel's arithmetic, re-scores it in fp64 and
 * then proves that no unselected tree can belong to the top `topk` (growing the short list when the proof
 * fails, up to BAMSE_TOPK_INTERNAL).  1 = the last call's result is proven identical to a full fp64 ranking;
 * 0 = the list hit the cap first (hundreds of trees tied within rounding): best effort.
 * bamse_last_topk_preselected: the preselection size the last call ended with. */
int bamse_last_topk_verified(const bamse_ctx* ctx);
int bamse_last_topk_preselected(const bamse_ctx* ctx);
/* Number of kernels this ctx has launched since creation (for bench.py `gpu_launches`). */
int64_t bamse_launch_count(const bamse_ctx* ctx);

/* ------------------------------------------------------------ problem upload
 * Replaces clustering.__init__'s table construction (clustering.py:56-62: As/Bs ->
 * K*M distrib() tables and the prior constant) for a batch of C clusterings and P
 * sequencing-error values.  var_counts/ref_counts are the aggregated per-cluster
 * variant/reference read counts `As`,`Bs` (clustering.py:56-57), laid out
 * [C][Kmax][M]; rows k >= K_of_c[c] are ignored.  Host work: the three lgamma terms of
 * distrib (clustering.py:23) in fp64; device work (kernel k_tables): the 512-point
 * tables.  The problem stays resident in HBM until the next upload / destroy. */
int bamse_upload_problem(bamse_ctx* ctx, int P, const double* err, int C, int M, int Kmax,
                         const int32_t* K_of_c, const int64_t* var_counts,
                         const int64_t* ref_counts);

/* Rebuilds every derived table of the resident problem from the read counts already in device memory
 * (what bamse_upload_problem does after its host->device copies): asynchronous on the ctx stream, no host
 * copies.  bench.py's device-resident `value` calls it every step, so that the table construction -- part of
 * the work per problem -- is inside the timed region. */
int bamse_rebuild_tables(bamse_ctx* ctx);

/* ------------------------------------------------------------------ score a list
 * Replaces clustering.get_tree_score(tree, nodes) (clustering.py:87-90) for a batch:
 * item i scores parent vector item_parent[i][0..k) (local indices, -1 at the root, any
 * root position, padding -2) over cluster ids item_nodes[i][0..k) (padding -1) of
 * clustering item_clust[i], for every parameter set p.  The prior always uses the
 * clustering's full K (clustering.py:61-62).  out_score is [P][n_items]: the raw
 * get_tree_score (no log scre).  Used by reorder (clustering.py:95-96), the greedy
 * candidate (:111) and the threshold (:121). */
int bamse_score_list(bamse_ctx* ctx, int64_t n_items, const int32_t* item_clust,
                     const int8_t* item_parent /* [n][Kmax] */,
                     const int8_t* item_nodes /* [n][Kmax] */, int precision,
                     double* out_score /* [P][n] */);

/* ------------------------------------------------------ score a dense index range
 * Scores trees [tree_begin, tree_begin + n) of clustering c under parameter set p by
 * index.  out_score[i] = get_tree_score, out_logscre[i] = log(canonizeF scre)
 * (tree_tools.py:47-63).  Either output may be NULL.  Parity-test / debugging entry. */
int bamse_score_range(bamse_ctx* ctx, int p, int c, uint64_t tree_begin, int64_t n,
                      int precision, double* out_score, double* out_logscre);

/* -------------------------------------------------------- enumerate + global top-k
 * Replaces the loop bamse.py:91-106 (get_trees2's exhaustive search space + score
 * assembly clustering.py:179-183 + argsort(-totalscore)[:t]).  For every p and every
 * clustering c, enumerates tree indices [tree_begin[c], tree_end[c]) (NULL, NULL = all
 * K_c^(K_c-1)), keeps trees with score + log scre >= threshold[p][c] (NULL = keep all;
 * clustering.py:144-147), ranks by total = score + log scre + clust_const[p][c]
 * (ties: clust asc, tree asc) and returns the best `topk` per p.
 *   In every precision a short list (topk + 8, doubled while the cut cannot be proven safe) is
 *   selected by the enumeration kernel with a small slack, re-scored by the fp64 list-scoring
 *   kernel before the exact threshold test and the final ranking: the returned totals/scores
 *   are fp64 check-mode values and BAMSE_FP32 / BAMSE_FP64 return the same records
 *   (see bamse_last_topk_verified).
 * out_records [P][topk] (unused slots: clust = -1), out_parent [P][topk][Kmax] parent
 * vectors (padding -2), out_count [P] number of valid records.  Host buffers. */
int bamse_score_topk(bamse_ctx* ctx, const double* clust_const /* [P][C] */,
                     const double* threshold /* [P][C] or NULL */,
                     const uint64_t* tree_begin /* [C] or NULL */,
                     const uint64_t* tree_end /* [C] or NULL */, int topk, int precision,
                     bamse_record* out_records, int8_t* out_parent, int32_t* out_count);

/* ------------------------------------------------- device-resident variants
 * Same enumeration, asynchronous on the ctx stream, no host copies: constants are read
 * from device memory and the per-p top `topk` records (selected in `precision`, no fp64
 * re-score) are written to d_out_records [P][topk].  d_threshold / d_tree_begin /
 * d_tree_end may be NULL.  `slack` is subtracted from thresholds (0 for exact). */
int bamse_enumerate_topk_dev(bamse_ctx* ctx, const double* d_clust_const,
                             const double* d_threshold, const uint64_t* d_tree_begin,
                             const uint64_t* d_tree_end, int topk, int precision,
                             double slack, bamse_record* d_out_records);

/* Merges n_lists device lists of [P][topk] records (e.g. the all-gathered per-rank
 * results) into one [P][topk] list with the same ordering rule.  Deterministic. */
int bamse_merge_topk_dev(bamse_ctx* ctx, const bamse_record* d_lists, int n_lists, int P,
                         int topk, bamse_record* d_out_records);

/* Number of trees the last bamse_enumerate_topk_dev / bamse_score_topk call on this ctx
 * scored (sum over p, c of the index-range sizes). */
int64_t bamse_last_eval_count(const bamse_ctx* ctx);

/* Device time of the enumeration kernel itself: every bamse_enumerate_topk_dev /
 * bamse_score_topk launch of it is bracketed by CUDA events on the ctx stream.  Returns
 * the summed elapsed milliseconds and the number of launches since the last reset
 * (synchronises the stream).  reset != 0 clears the accumulator after reading. */
int bamse_kernel_time(bamse_ctx* ctx, int reset, double* out_ms, int64_t* out_launches);

/* The same for the table-building kernels: every bamse_upload_problem / bamse_rebuild_tables brackets its launches
 * (binomial tables, inside levels, context levels) by CUDA events; summed milliseconds and number of builds. */
int bamse_table_time(bamse_ctx* ctx, int reset, double* out_ms, int64_t* out_builds);

/* Executed-work counters of the fp32 enumeration kernel since the last reset:
 * out4 = { MUFU (ex2/lg2) thread-level ops, convolution terms exponentiated,
 *          convolutions, prefix scans }.  For the pipe-utilisation side of the roofline. */
int bamse_kernel_stats(bamse_ctx* ctx, int reset, uint64_t* out4);

/* ---------------------------------------------------------- introspection (no GPU needed)
 * The fp32 path memoises sub-forest messages in tables and drives each tree by a small program
 * (bamse_b200/csrc/forest.cuh).  Both are data independent; these entry points run the HOST instances of
 * the code the kernels run, so tests can interpret them with the CPU oracle.
 * bamse_debug_forest_geometry: *s / *sh > 0 on entry = use that geometry, else the default of K is returned;
 *   out_nf / out_nfh = table rows per (parameter set, clustering, sample) / per root; out_slots = live vectors.
 * bamse_debug_forest_recipes: out [nf][5] = {kind, root label, row a, row b, node mask} of every table row.
 * bamse_debug_program: the program of a tree (k nodes, parent vector, optional cluster ids) in a clustering
 *   of Kc clusters: out = {n_ops, fin, fin_arg, fin_r, scre, max_sp, op[20], arg[20]}. */
int bamse_debug_forest_geometry(int K, int32_t* s, int32_t* sh, int32_t* out_nf, int32_t* out_nfh, int32_t* out_slots);
int bamse_debug_forest_recipes(int K, int s, int32_t* out);
int bamse_debug_program(int k, const int8_t* parent, const int8_t* nodes, int Kc, int s, int sh, int32_t* out);
/* Inside x outside pairs (bamse_b200/csrc/pairs.cuh): a tree whose best cut leaves a pack of <= s nodes and a
 * context of <= r nodes is scored by one dot product of an inside-table row with a context-table row.
 * bamse_debug_pair_geometry: *s <= 0 on entry = the default (s, r) of K is returned; out_nf / out_nt = inside /
 *   context rows per (parameter set, clustering, sample).
 * bamse_debug_ctx_recipes: out [nt][3] = {context row of the marked node's parent (-1: the marked node is the
 *   root), inside row of the marked node's children forest within the context (-1: none), marked label}.
 * bamse_debug_classify: the cut of a full tree on K labels: out = {regular (0 / 1), inside row, context row}. */
int bamse_debug_pair_geometry(int K, int32_t* s, int32_t* r, int32_t* out_nf, int32_t* out_nt);
int bamse_debug_ctx_recipes(int K, int s, int r, int32_t* out);
int bamse_debug_classify(int K, const int8_t* parent, int s, int r, int32_t* out);
/* The assignment step of the sharding plan as a pure function (no GPU): build_cost / enum_cost are indexed by cluster
 * count [0 .. BAMSE_KMAX]; out_owner [P][C] = owning rank or -1 (split over all ranks). */
int bamse_debug_shard_plan(int P, int C, const int32_t* K_of_c, int world, const double* build_cost,
                           const double* enum_cost, int32_t* out_owner);
/* Needs the GPU (cuts all K^(K-1) trees on the device): out4 = {inside rows, context rows of the full index space,
 * context rows in use, irregular trees} of the pair geometry (s, r).  Invalidates the resident problem. */
int bamse_debug_pair_stats(bamse_ctx* ctx, int K, int s, int r, int64_t* out4);

/* Host helper shared with the oracle's definition: index -> parent vector. */
int bamse_decode_tree(int K, uint64_t tree_index, int8_t* out_parent /* [K] */);

#ifdef __cplusplus
}
#endif
#endif /* BAMSE_CUDA_H */
