// Kernel declarations shared between the kernel translation units and the C ABI.
#pragma once
#include "common.cuh"
#include "forest.cuh"
#include "pairs.cuh"

namespace bamse {

// One explicit work item of bamse_score_list / the fp64 re-score of a short list.
struct ScoreItem {
    int32_t p;
    int32_t c;
    int8_t k;
    int8_t parent[KMAX];
    int8_t nodes[KMAX];
    int8_t pad[3];
};
static_assert(sizeof(ScoreItem) == 36, "ScoreItem layout");

// One (parameter set, clustering, tree-index range) segment of an enumeration.
struct Segment {
    int32_t p;
    int32_t c;
    int32_t K;
    int32_t stride;          // tree index step (interleaved sharding: rank r scores begin + r + i * world)
    uint64_t tree_begin;
    uint64_t n_trees;
    uint64_t first_chunk;    // prefix sum of chunk counts
    double clust_const;      // cluster_prior + maxlikelihood        (clustering.py:183)
    double threshold;        // candidate score + log scre (minus slack), -inf = keep all
    int32_t whole;           // 1: the segment is the complete index range [0, K^(K-1)) of its clustering
    int32_t pad;
};

struct EnumWork {
    const Segment* segs;     // [n_segs]
    int n_segs;
    int chunk;               // trees per work chunk
    const uint64_t* total_chunks;     // device scalar written by k_build_segments
    unsigned long long* next_chunk;   // device work counter (zeroed before launch)
    int topk;
    bamse_record* cand;      // flat bag of candidate records appended by the workers (each carries its param)
    unsigned long long* cand_count;   // append cursor (zeroed before launch)
    unsigned long long cand_cap;      // capacity in records (sized so that it cannot overflow)
    double* dense_score;     // optional [n_trees of segment 0]
    double* dense_logscre;   // optional
    unsigned long long* stats;   // optional [4]: executed MUFU ops (thread level), conv terms, convs, scans
    unsigned long long* samples_done;   // optional: tree-samples actually evaluated (early exit statistics)
    int early_exit;              // fp32 kernel: stop a tree once its partial sum over samples cannot reach the cut
    // tail splitting (fp32 kernel, chunk == 1): the last `n_split` trees of the launch are issued as
    // split_ways work items each (samples m = h, h + ways, ...), combined through split_acc/split_cnt
    int split_ways;              // 1 = off
    unsigned long long n_split;  // upper bound on the trees issued split (size of the two buffers)
    double* split_acc;           // [n_split] partial sums (zeroed before launch)
    int* split_cnt;              // [n_split] finished parts (zeroed before launch)
    // fp32 kernel: per-K tables of ready-made evaluation programs, indexed by tree index (device array of
    // KMAX + 1 device pointers, entry NULL = decode in the kernel); NULL = no tables at all
    const FastProgram* const* prog_tabs;
    // fp32 kernel: the workers' private top-k lists, [workers][max(topk, 1)] records in global memory
    // (touched once per tree by one lane; keeping them out of shared memory buys a resident CTA per SM)
    bamse_record* lists;
    // pair path (pairs.cuh): k_enumerate_pairs scores the regular trees of every segment (one lane per tree) and
    // appends the irregular ones, as (segment << 40) | tree, to the region of their parameter set in `flat`;
    // k_flat_prefix turns the per-p counts into running sums, and k_enumerate_fast in flat mode (flat_mode = 1)
    // takes its work from that list instead of the segments' index ranges
    unsigned long long* flat;             // NULL: no pair path
    const unsigned long long* flat_off;   // [P + 1] first slot of every parameter set's region
    unsigned long long* flat_cnt;         // [P] entries appended (zeroed before launch)
    unsigned long long* flat_prefix;      // [P + 1] running sums of flat_cnt
    unsigned long long* flat_overflow;    // set when a region was too small (host checks)
    int flat_mode;
    bamse_record* lists2;                 // the pair kernel's per-warp lists
    float* part;                          // sample-major pair path: per-sample log2 results [M][tree slots]
    unsigned long long* trees_done;       // trees this rank took in k_enumerate_pairs (regular + irregular)
    int grab;                             // k_enumerate_pairs: chunks taken per visit of the work counter
};

void launch_tables(const ProblemView& pv, const double* d_var, const double* d_ref, const double* d_logc,
                   const double* d_logcf, const double* d_log1mcf, double* D64, float* D32,
                   cudaStream_t s);
void launch_build_segments(int P, int C, const int32_t* d_K_of_c, const double* d_const, const double* d_thr,
                           const uint64_t* d_begin, const uint64_t* d_end, double slack_abs, double slack_rel,
                           int chunk, int reps /* chunks per `chunk` trees */, int shard_rank, int shard_world,
                           uint32_t k_mask /* cluster counts that take part */, int split_full /* 1: a (p, c) split over
                           the ranks gets its whole range on every rank (node-set sharding inside the pair kernel) */,
                           const int32_t* d_owner /* [P][C] sharding plan or NULL */, Segment* d_segs,
                           uint64_t* d_total_chunks, uint64_t* d_total_trees, cudaStream_t s);
int enumerate_grid_size(int precision, int device);
void launch_enumerate(int precision, const ProblemView& pv, const EnumWork& w, int grid, cudaStream_t s);
void launch_score_items(int precision, const ProblemView& pv, const ScoreItem* d_items, int64_t n,
                        double* d_out, cudaStream_t s);
// d_bag: n records (min(*d_count, n) when d_count is non-NULL) in any order, each tagged with its param;
// writes the best topk of every parameter set to d_out [P][topk]
void launch_merge_topk(const bamse_record* d_bag, const unsigned long long* d_count, unsigned long long n, int P,
                       int topk, bamse_record* d_out, cudaStream_t s);
int enumerate_workers(int precision, int grid);

struct FastTables;
void launch_tables_fast(const ProblemView& pv, const FastTables& ft, float* D2, float* FG, float* FSG, float* SD2,
                        cudaStream_t s);
// level n of the message tables (max_rows = the largest number of n-node forests of any clustering present)
// fsg_top = 0: the scanned messages of the top level (n == s) are left to launch_forest_scan_top
void launch_forest_level(const ProblemView& pv, const FastTables& ft, int n, int max_rows, float* FG, float* FSG,
                         int fsg_top, unsigned long long* stats /* NULL: no executed-work counters */, cudaStream_t s);
void launch_forest_scan_top(const ProblemView& pv, const FastTables& ft, int max_rows, const float* FG, float* FSG,
                            cudaStream_t s);
// finishing tables (max_items = the largest K * nfh of any clustering present)
void launch_forest_fh(const ProblemView& pv, const FastTables& ft, int max_items, float* FH, cudaStream_t s);
int enumerate_fast_grid_size(int device, int topk, bool framed, int slots_needed);
// out[i] = evaluation program of tree index i of a fi.K-node clustering, i < n_trees = K^(K-1)
void launch_build_prog_table(const ForestIndex& fi, uint64_t n_trees, const int16_t* d_lut, FastProgram* out, cudaStream_t s);
void launch_enumerate_fast(const ProblemView& pv, const FastTables& ft, const EnumWork& w, int grid, bool framed,
                           int slots_needed, cudaStream_t s);
// context tables, level q (max_rows = the largest number of q-node contexts of any cluster count present)
void launch_ctx_level(const ProblemView& pv, const FastTables& ft, int q, int max_rows, float* T, unsigned long long* stats,
                      cudaStream_t s);
// pair table of a cluster count, two passes: out == NULL: used[row] = 1 for every context row (FULL index space)
// some regular tree refers to, *n_irregular += irregular trees; out != NULL: out[tree] = pair_entry with the
// compact context row d_map[row]
void launch_build_pair_table(const ForestIndex& fi, const PairGeom& g, uint64_t n_trees, const int16_t* d_lut,
                             uint64_t* out, uint8_t* used, unsigned long long* n_irregular, const uint32_t* d_map,
                             const ForestIndex& fi_prog /* geometry of the programs */, FastProgram* irr_progs /* NULL: none;
                             else the program of every irregular tree, slot = entry & (2^42 - 1), *n_irregular = slot cursor */,
                             unsigned long long* mask_hist /* NULL or [2^K]: regular trees per pack node set */, cudaStream_t s);
int enumerate_pairs_grid_size(int device);
int enumerate_pairs_workers(int grid);
// mode 0: tree-major; 1 + 2: sample-major (accumulate per sample into w.part, then sum and rank); the segments of
// modes 1 / 2 carry M chunks per 32 trees (launch_build_segments with reps = M)
void launch_enumerate_pairs(const ProblemView& pv, const FastTables& ft, const EnumWork& w, int grid, int mode, cudaStream_t s);
void launch_pair_hist(const uint64_t* tab, uint64_t n_trees, unsigned* hist, cudaStream_t s);
void launch_pair_scatter(const uint64_t* tab, uint64_t n_trees, unsigned* cursor, uint32_t* sorted_tree, uint64_t* sorted_entry,
                         cudaStream_t s);
void launch_flat_prefix(int P, const unsigned long long* cnt, const unsigned long long* off, unsigned long long* prefix,
                        cudaStream_t s);
void launch_score_items_fast(const ProblemView& pv, const FastTables& ft, const ScoreItem* d_items, int64_t n,
                             double* d_out, cudaStream_t s);

}  // namespace bamse
