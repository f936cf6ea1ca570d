// C ABI of libbamse_cuda.so (include/bamse_cuda.h).  Host-side orchestration only:
// argument checks, host<->device copies, kernel launches.  No CPU scoring path exists.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "score_fast.cuh"
#include "tree_decode.cuh"
#include "forest_host.h"
#include "pairs_host.h"

using namespace bamse;

namespace {
thread_local std::string g_create_error;

struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
}  // namespace

struct bamse_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int64_t last_evals = 0;

    // resident problem
    bool have_problem = false;
    int P = 0, C = 0, M = 0, Kmax = 0;
    std::vector<int32_t> K_of_c;
    DBuf d_K, d_p0, d_var, d_ref, d_logc, d_logcf, d_log1mcf, d_D64, d_D32, d_D2, d_SD2;
    // sub-forest tables (forest.cuh): message tables FG / FSG, finishing tables FH, per-(p, c) bases, shape LUT,
    // per-K recipes; geometry (s, sh) per cluster count of the resident problem
    DBuf d_FG, d_FSG, d_FH, d_tb, d_lut, d_recipe_ptrs, d_order_ptrs;
    struct Recipes { DBuf buf, order; int s = -1; };
    Recipes recipes[KMAX + 1];
    int geom_s[KMAX + 1] = {0}, geom_sh[KMAX + 1] = {0}, geom_r[KMAX + 1] = {0};
    // pair path (pairs.cuh), per cluster count: pair table, context map / compact recipes; kept across uploads
    // while the geometry of that K is unchanged
    struct PairState {
        DBuf tab, map, recipes, sorted_tree, sorted_entry, irr_progs, mask_owner;
        std::vector<unsigned long long> mask_trees;      // regular trees per pack node set (data independent)
        int owner_world = -1;                            // the world size mask_owner was balanced for
        int s = -1, r = -1, sh = -1;
        bool has_tab = false;
        uint32_t nt = 0, lbase[PR_RMAX + 2] = {0};
        uint64_t n_irregular = 0;
    };
    PairState pairs[KMAX + 1];
    DBuf d_T, d_pk, d_flat, d_flat_meta, d_lists2, d_segs2, d_owner, d_part, d_pcm;
    int n_pcm = 0;
    std::vector<int32_t> owner;          // [P][C] sharding plan (empty: none): -1 split over all ranks, else the owning rank
    bool k_present[KMAX + 1] = {false};
    bool interp_ready = false;           // finishing tables + top-level scanned messages of the resident problem built
    int max_top_rows = 0, max_fh_items = 0, max_level_rows[FT_SMAX + 2] = {0}, max_ctx_rows[PR_RMAX + 2] = {0};
    int smax = 1, rmax = 0;
    size_t fh_rows = 0;
    int grid_pairs = 0;
    int slots_needed = 3;
    double table_bytes = 0.0;
    // program tables (fp32 enumeration): one per K <= PROGTAB_KMAX present in the problem, kept across
    // uploads while the table geometry of that K is unchanged
    struct ProgTab { DBuf buf; int s = -1, sh = -1; };
    ProgTab progtab[KMAX + 1];
    DBuf d_progtabs;         // [KMAX + 1] device pointers, NULL = decode in the kernel
    bool have_progtabs = false;

    // work buffers
    DBuf d_segs, d_scalars /* total_chunks, total_trees, next_chunk */, d_cand, d_items, d_scores;
    DBuf d_const, d_thr, d_begin, d_end, d_records, d_dense0, d_dense1, d_split, d_lists;
    int grid[3] = {0, 0, 0};
    uint64_t last_cand_cap = 0;   // capacity of the candidate bag of the last enumeration
    int shard_rank = 0, shard_world = 1;
    int plan_rank = 0, plan_world = 1;   // the shard the resident problem's plan / tables were made for
    std::vector<int32_t> plan_shape;     // what plan_tables was last run for (problem shape + shard + switches)
    bool plan_valid = false;
    int shard_plan = 1;      // BAMSE_OPT_SHARD_PLAN
    int collect_stats = 0;   // BAMSE_OPT_STATS
    int early_exit = 0;      // BAMSE_OPT_EARLY_EXIT
    int last_verified = 1;   // the last bamse_score_topk proved its preselection cut safe (bamse_last_topk_verified)
    int last_ksel = 0;       // ... and the preselection size it ended with
    int use_framed = -1;     // fp32 kernel variant for the resident problem: -1 not probed yet, 0 exponential, 1 framed
    bool probing = false;
    uint64_t problem_sig = 0, probed_sig = 0;   // the variant choice is memoised per problem content
    int probed_choice = -1;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_used, ev_build_used, ev_free;

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    FastTables fast() const {
        FastTables f;
        f.D2 = d_D2.as<float>(); f.SD2 = d_SD2.as<float>();
        f.FG = d_FG.as<float>(); f.FSG = d_FSG.as<float>(); f.FH = d_FH.as<float>();
        f.T = d_T.as<float>(); f.pk = d_pk.as<PairK>();
        f.pcm_list = d_pcm.as<int32_t>(); f.n_pcm = n_pcm; f.grid_blocks = 1;
        f.shard_rank = plan_rank; f.shard_world = plan_world;
        f.tb = d_tb.as<TableBase>(); f.lut = d_lut.as<int16_t>();
        f.recipes = d_recipe_ptrs.as<const ForestRecipe*>();
        f.level_order = d_order_ptrs.as<const uint32_t*>();
        return f;
    }
    ProblemView view() const {
        ProblemView v;
        v.P = P; v.C = C; v.M = M; v.Kmax = Kmax;
        v.K_of_c = d_K.as<int32_t>(); v.p0 = d_p0.as<double>();
        v.D64 = d_D64.as<double>(); v.D32 = d_D32.as<float>();
        return v;
    }
};

constexpr int PROGTAB_KMAX = 8;      // K = 8: 2 M programs, 134 MB
constexpr int PAIRTAB_KMAX = 10;     // K = 10: 10^9 entries, 8 GB (kept across uploads); above, the cut is computed in the kernel
constexpr int TOPK_INTERNAL = BAMSE_TOPK_INTERNAL;   // preselection lists may grow to this size (public topk <= BAMSE_TOPK_MAX)

// A C++ exception (std::bad_alloc of a staging vector, in practice) must not cross the C ABI: the entry
// points that allocate host memory are function-try-blocks that end here.
static int on_exception(bamse_ctx* ctx, const char* fn) noexcept {
    int code = BAMSE_EINTERNAL;
    const char* what = "unexpected C++ exception";
    try { throw; }
    catch (const std::bad_alloc&) { code = BAMSE_ENOMEM; what = "out of host memory"; }
    catch (const std::exception&) {}
    catch (...) {}
    try { if (ctx) ctx->fail(code, "%s: %s", fn, what); } catch (...) {}
    return code;
}

static inline bool valid_precision(int p) { return p == BAMSE_FP32 || p == BAMSE_FP64 || p == BAMSE_FP32_LOGDOMAIN; }

#define CTX_CHECK(ctx)                      \
    do {                                    \
        if (!(ctx)) return BAMSE_EINVAL;    \
        cudaSetDevice((ctx)->device);       \
    } while (0)


static bool pairs_enabled() {
    const char* e = getenv("BAMSE_PAIRS");           // BAMSE_PAIRS=0: table-driven programs only (A/B measurements)
    return !(e && atoi(e) == 0);
}

// Pair state of a cluster count for the geometry (geom_s, geom_r): the cut of every tree (pair table for
// K <= PAIRTAB_KMAX, a marking pass above), the context rows in use closed under the recipes' parent links, the
// map from the full context index space to compact rows and the compact recipes.  Data independent: built once
// per ctx and geometry.
static int ensure_pair_state(bamse_ctx* ctx, int K) {
    bamse_ctx::PairState& ps = ctx->pairs[K];
    const int sK = ctx->geom_s[K], rK = ctx->geom_r[K], shK = ctx->geom_sh[K];
    if (ps.s == sK && ps.r == rK && ps.sh == shK) return BAMSE_OK;
    ps.s = ps.r = ps.sh = -1;
    const ForestIndex fi = make_forest_index(K, sK, 1);
    const PairGeom g = make_pair_geom(K, sK, rK);
    if (fi.nf >= (1 << PAIR_FG_BITS)) return ctx->fail(BAMSE_EINTERNAL, "pair geometry of K=%d exceeds %u-bit inside row ids", K, PAIR_FG_BITS);
    cudaStream_t st = ctx->stream;
    const uint64_t n_trees = ipow_u64(K, K - 1);
    ps.has_tab = K <= PAIRTAB_KMAX;
    if (const char* e = getenv("BAMSE_PAIRTAB_KMAX")) ps.has_tab = K <= atoi(e);
    DBuf used;
    const size_t used_bytes = ((size_t)g.nt + 15) / 8 * 8;
    cudaError_t e = used.ensure(used_bytes + 8);
    if (e != cudaSuccess) return ctx->fail(BAMSE_ENOMEM, "pair state: %s", cudaGetErrorString(e));
    cudaMemsetAsync(used.p, 0, used_bytes + 8, st);
    unsigned long long* d_nirr = reinterpret_cast<unsigned long long*>(used.as<uint8_t>() + used_bytes);
    launch_build_pair_table(fi, g, n_trees, ctx->d_lut.as<int16_t>(), nullptr, used.as<uint8_t>(), d_nirr, nullptr, fi, nullptr, nullptr, st);
    ctx->launches++;
    std::vector<uint8_t> h_used(used_bytes + 8);
    e = cudaMemcpyAsync(h_used.data(), used.p, used_bytes + 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { used.release(); return ctx->fail(BAMSE_ECUDA, "pair state of K=%d: %s", K, cudaGetErrorString(e)); }
    unsigned long long n_irr = 0;
    memcpy(&n_irr, h_used.data() + used_bytes, sizeof n_irr);
    used.release();
    // recipes of the rows in use only (K = 10: 2.4 M of 10.6 M), level by level from the largest contexts down: a
    // row in use needs the context its recipe starts from, which has fewer nodes; worker threads split each level
    std::vector<CtxRecipe> full((size_t)g.nt);
    {
        unsigned n_thr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        for (int q = rK; q >= 1; --q) {
            const uint32_t lo = g.base[q], hi = g.base[q + 1];
            if (hi - lo < 20000) n_thr = 1;
            auto work = [&](uint32_t a, uint32_t b) {
                for (uint32_t i = a; i < b; ++i) if (h_used[i]) full[i] = ctx_recipe_of_row(fi, g, i);
            };
            if (n_thr == 1) work(lo, hi);
            else {
                std::vector<std::thread> pool;
                const uint32_t step = (hi - lo + n_thr - 1) / n_thr;
                for (unsigned t = 0; t < n_thr; ++t) {
                    const uint32_t a = lo + t * step, b = std::min(hi, a + step);
                    if (a < b) pool.emplace_back(work, a, b);
                }
                for (auto& th : pool) th.join();
            }
            for (uint32_t i = lo; i < hi; ++i)
                if (h_used[i] && full[i].parent_row >= 0) h_used[(size_t)full[i].parent_row] = 1;
        }
    }
    std::vector<uint32_t> map(g.nt, 0xffffffffu);
    std::vector<CtxRecipe> comp;
    for (int q = 1; q <= PR_RMAX + 1; ++q) {
        ps.lbase[q] = (uint32_t)comp.size();
        if (q > rK) continue;
        for (int pass = 0; pass < 2; ++pass)          // rows that need a correlation first: CTAs of one kind
            for (uint32_t i = g.base[q]; i < g.base[q + 1]; ++i) {
                if (!h_used[i] || (full[i].h_row >= 0) != (pass == 0)) continue;
                map[i] = (uint32_t)comp.size();
                CtxRecipe c = full[i];
                if (c.parent_row >= 0) c.parent_row = (int32_t)map[(size_t)c.parent_row];
                comp.push_back(c);
            }
    }
    ps.lbase[0] = 0;
    ps.nt = (uint32_t)comp.size();
    BAMSE_CUDA_TRY(ctx, ps.map.ensure(sizeof(uint32_t) * std::max<size_t>(map.size(), 1)));
    BAMSE_CUDA_TRY(ctx, ps.recipes.ensure(sizeof(CtxRecipe) * std::max<size_t>(comp.size(), 1)));
    // (copies on the ctx stream: a blocking cudaMemcpy from pageable memory may return before the DMA has landed
    //  and is not ordered against this non-blocking stream)
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ps.map.p, map.data(), sizeof(uint32_t) * map.size(), cudaMemcpyHostToDevice, st));
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ps.recipes.p, comp.data(), sizeof(CtxRecipe) * comp.size(), cudaMemcpyHostToDevice, st));
    if (ps.nt >= (1u << PAIR_T_BITS)) return ctx->fail(BAMSE_EINTERNAL, "pair geometry of K=%d needs more than %u-bit context row ids", K, PAIR_T_BITS);
    if (ps.has_tab) {
        if (ps.tab.ensure(sizeof(uint64_t) * n_trees) != cudaSuccess) { cudaGetLastError(); ps.has_tab = false; }   // no room: cut in the kernel
        else {
            // the irregular trees' programs with the table (K = 10: 20 M x 64 bytes), unless they are too many
            ps.irr_progs.release();
            DBuf cursor;
            const ForestIndex fi_prog = make_forest_index(K, program_forest_s(K, sK), shK);
            const bool want_progs = n_irr > 0 && n_irr * sizeof(FastProgram) <= (6ull << 30) &&
                                    cursor.ensure(sizeof(unsigned long long)) == cudaSuccess &&
                                    ps.irr_progs.ensure(sizeof(FastProgram) * n_irr) == cudaSuccess;
            if (want_progs) cudaMemsetAsync(cursor.p, 0, sizeof(unsigned long long), st);
            else { cudaGetLastError(); ps.irr_progs.release(); }
            DBuf mhist;                                 // regular trees per pack node set, for the node-set sharding
            const size_t n_masks = (size_t)1 << K;
            const bool want_hist = mhist.ensure(sizeof(unsigned long long) * n_masks) == cudaSuccess;
            if (want_hist) cudaMemsetAsync(mhist.p, 0, sizeof(unsigned long long) * n_masks, st);
            launch_build_pair_table(fi, g, n_trees, ctx->d_lut.as<int16_t>(), ps.tab.as<uint64_t>(), nullptr,
                                    want_progs ? cursor.as<unsigned long long>() : nullptr, ps.map.as<uint32_t>(), fi_prog,
                                    want_progs ? ps.irr_progs.as<FastProgram>() : nullptr,
                                    want_hist ? mhist.as<unsigned long long>() : nullptr, st);
            ctx->launches++;
            ps.mask_trees.assign(want_hist ? n_masks : 0, 0ull);
            if (want_hist) BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ps.mask_trees.data(), mhist.p, sizeof(unsigned long long) * n_masks, cudaMemcpyDeviceToHost, st));
            BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(st));
            ps.owner_world = -1;
            cursor.release(); mhist.release();
        }
    }
    // the same table in the order of the inside rows (counting sort; BAMSE_PAIR_SORT=0 disables)
    ps.sorted_tree.release(); ps.sorted_entry.release();
    const char* no_sort = getenv("BAMSE_PAIR_SORT");
    if (ps.has_tab && n_trees < (1ull << 32) && !(no_sort && atoi(no_sort) == 0)) {
        DBuf hist;
        bool ok = hist.ensure(sizeof(unsigned) * (size_t)fi.nf) == cudaSuccess &&
                  ps.sorted_tree.ensure(sizeof(uint32_t) * n_trees) == cudaSuccess &&
                  ps.sorted_entry.ensure(sizeof(uint64_t) * n_trees) == cudaSuccess;
        if (ok) {
            cudaMemsetAsync(hist.p, 0, sizeof(unsigned) * (size_t)fi.nf, st);
            launch_pair_hist(ps.tab.as<uint64_t>(), n_trees, hist.as<unsigned>(), st);
            std::vector<unsigned> h((size_t)fi.nf);
            ok = cudaMemcpyAsync(h.data(), hist.p, sizeof(unsigned) * h.size(), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                 cudaStreamSynchronize(st) == cudaSuccess;
            if (ok) {
                uint64_t acc = 0;
                for (size_t i = 0; i < h.size(); ++i) { const unsigned c = h[i]; h[i] = (unsigned)acc; acc += c; }
                ok = acc == n_trees && cudaMemcpyAsync(hist.p, h.data(), sizeof(unsigned) * h.size(), cudaMemcpyHostToDevice, st) == cudaSuccess;
            }
            if (ok) {
                launch_pair_scatter(ps.tab.as<uint64_t>(), n_trees, hist.as<unsigned>(), ps.sorted_tree.as<uint32_t>(),
                                    ps.sorted_entry.as<uint64_t>(), st);
                ok = cudaStreamSynchronize(st) == cudaSuccess;
                ctx->launches += 2;
            }
        }
        hist.release();
        if (!ok) { cudaGetLastError(); ps.sorted_tree.release(); ps.sorted_entry.release(); }   // no room: index order
    }
    BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(st));   // the host vectors above die at return
    BAMSE_CUDA_TRY(ctx, cudaGetLastError());
    ps.n_irregular = n_irr;
    ps.s = sK; ps.r = rK; ps.sh = shK;
    return BAMSE_OK;
}

// The assignment itself (pure host arithmetic; bamse_debug_shard_plan exposes it to the CPU tests): longest
// processing time first over whole (p, c) pairs, with the j most expensive pairs split over all ranks instead (a
// split pair costs its table build on every rank and 1/W of its enumeration), j chosen to minimise the makespan.
static std::vector<int32_t> plan_owners(int P, int C, const int32_t* K_of_c, int W, const double* build_k, const double* enum_k) {
    const int n = P * C;
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    auto cost = [&](int i) { const int K = K_of_c[i % C]; return build_k[K] + enum_k[K]; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
    int best_j = 0;
    double best_span = -1.0;
    std::vector<double> load(W);
    const int j_max = std::min(n, 4 * W + 8);
    for (int j = 0; j <= j_max; ++j) {
        double base = 0.0;
        for (int t = 0; t < j; ++t) { const int K = K_of_c[order[t] % C]; base += build_k[K] + enum_k[K] / W; }
        for (int r = 0; r < W; ++r) load[r] = base;
        for (int t = j; t < n; ++t) {
            int r_min = 0;
            for (int r = 1; r < W; ++r) if (load[r] < load[r_min]) r_min = r;
            load[r_min] += cost(order[t]);
        }
        double span = 0.0;
        for (int r = 0; r < W; ++r) span = std::max(span, load[r]);
        if (best_span < 0.0 || span < best_span * (1.0 - 1e-9)) { best_span = span; best_j = j; }
    }
    std::vector<int32_t> owner(n, -1);
    for (int r = 0; r < W; ++r) load[r] = 0.0;
    for (int t = best_j; t < n; ++t) {
        int r_min = 0;
        for (int r = 1; r < W; ++r) if (load[r] < load[r_min]) r_min = r;
        load[r_min] += cost(order[t]);
        owner[order[t]] = r_min;
    }
    return owner;
}

// Sharding plan (multi-GPU): which rank scores which (parameter set, clustering).  Tables are per (p, c), so a
// pair that one rank owns entirely costs its table build once; a pair split over all ranks (interleaved tree
// indices) costs the build on EVERY rank.  Longest-processing-time assignment of whole pairs, with the j most
// expensive pairs split instead, j chosen to minimise the estimated makespan (small problems: a few whole pairs
// per rank; one huge clustering: split).  Costs are in convolution units per sample: table rows that need a
// convolution count 1, the other rows 0.25, a regular tree 0.004 (one windowed dot product), an irregular tree or
// a tree of a cluster count without pair path ~1.  Deterministic: every rank computes the same plan.
static void make_shard_plan(bamse_ctx* ctx) {
    ctx->owner.clear();
    ctx->plan_rank = ctx->shard_rank; ctx->plan_world = ctx->shard_world;
    const int W = ctx->shard_world, P = ctx->P, C = ctx->C;
    if (W <= 1 || !ctx->shard_plan) return;
    if (const char* e = getenv("BAMSE_SHARD_PLAN")) if (atoi(e) == 0) return;
    double build_k[KMAX + 1] = {0.0}, enum_k[KMAX + 1] = {0.0};
    for (int K = 1; K <= KMAX; ++K) {
        if (!ctx->k_present[K]) continue;
        const ForestIndex fi = make_forest_index(K, ctx->geom_s[K], ctx->geom_sh[K]);
        double conv_rows = 0.0;
        for (int n = 2; n <= fi.s; ++n) {
            double trees = 1.0;                       // n^(n-1) single trees among the (n+1)^(n-1) forests
            for (int i = 0; i < n - 1; ++i) trees *= n;
            conv_rows += binom_small(K, n) * ((double)forest_count(n) - trees);
        }
        const double T = (double)ipow_u64(K, K - 1);
        if (ctx->geom_r[K] > 0) {
            build_k[K] = conv_rows + 0.25 * fi.nf + 1.1 * ctx->pairs[K].nt;
            enum_k[K] = 0.004 * T + 1.2 * (double)ctx->pairs[K].n_irregular;
        } else {
            build_k[K] = conv_rows + 0.25 * fi.nf + (double)K * fi.nfh;
            enum_k[K] = T * (K <= 6 ? 0.1 : 0.25 * (K - 5));
        }
    }
    ctx->owner = plan_owners(P, C, ctx->K_of_c.data(), W, build_k, enum_k);
}

// Geometry of every cluster count present (stepped down until the tables fit the budget), pair states, table
// bases, allocations, recipes and program tables of the resident problem.  Host work + data-independent kernels.
static int plan_tables(bamse_ctx* ctx) {
    const int P = ctx->P, C = ctx->C, M = ctx->M;
    const bool* k_present = ctx->k_present;
    cudaStream_t s = ctx->stream;
    if (!ctx->d_lut.p) {
        BAMSE_CUDA_TRY(ctx, ctx->d_lut.ensure(sizeof(int16_t) * FT_LUT_SIZE));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_lut.p, forest_lut().lut.data(), sizeof(int16_t) * FT_LUT_SIZE, cudaMemcpyHostToDevice, s));
    }
    // BAMSE_TABLE_GB: table budget (default 60 % of the device memory); BAMSE_FOREST_S / _SH / BAMSE_PAIR_R override
    const char* env_s = getenv("BAMSE_FOREST_S");
    const char* env_sh = getenv("BAMSE_FOREST_SH");
    const char* env_r = getenv("BAMSE_PAIR_R");
    const bool pairs_on = pairs_enabled();
    for (int K = 1; K <= KMAX; ++K) {
        if (!k_present[K]) continue;
        default_forest_geometry(K, &ctx->geom_s[K], &ctx->geom_sh[K], &ctx->geom_r[K]);
        if (env_s) ctx->geom_s[K] = std::max(1, std::min(atoi(env_s), FT_SMAX));
        if (env_sh) ctx->geom_sh[K] = std::max(1, atoi(env_sh));
        if (env_r) ctx->geom_r[K] = std::max(0, std::min(atoi(env_r), PR_RMAX));
        int& sK = ctx->geom_s[K];
        int& rK = ctx->geom_r[K];
        if (!pairs_on || K < 2 || sK + rK < K || rK - 1 > sK) rK = 0;
        if (rK > 0 && make_forest_index(K, sK, 1).nf >= (1 << PAIR_FG_BITS)) rK = 0;
        if (rK == 0) sK = program_forest_s(K, sK);     // programs only: nothing above the 16-bit row ids is needed
        ctx->geom_sh[K] = std::min(ctx->geom_sh[K], program_forest_s(K, sK));
    }
    double budget = 0.0;
    std::vector<TableBase> tbs((size_t)P * C);
    size_t fg_rows = 0, fsg_rows = 0, fh_rows = 0, t_rows = 0;
    for (;;) {
        for (int K = 1; K <= KMAX; ++K)
            if (k_present[K] && ctx->geom_r[K] > 0) { const int rc = ensure_pair_state(ctx, K); if (rc) return rc; }
        {   // 60 % of the device, but never more than what is free once the data-independent tables of the cluster
            // counts present (pair tables: 21 GB at K = 10) are in place; the table buffers held now are re-used
            size_t free_b = 0, total_b = 0;
            BAMSE_CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
            const double held = (double)(ctx->d_FG.cap + ctx->d_FSG.cap + ctx->d_T.cap + ctx->d_FH.cap);
            budget = std::min(0.6 * (double)total_b, 0.9 * ((double)free_b + held));
            if (const char* g = getenv("BAMSE_TABLE_GB")) budget = atof(g) * 1e9;
        }
        make_shard_plan(ctx);
        fg_rows = fsg_rows = fh_rows = t_rows = 0;
        double per_k[KMAX + 1] = {0.0};
        for (int p = 0; p < P; ++p)
            for (int c = 0; c < C; ++c) {
                const int K = ctx->K_of_c[c];
                TableBase& t = tbs[(size_t)p * C + c];
                t.fi = make_forest_index(K, ctx->geom_s[K], ctx->geom_sh[K]);
                t.s_prog = program_forest_s(K, ctx->geom_s[K]);
                t.mine = ctx->owner.empty() || ctx->owner[(size_t)p * C + c] < 0 || ctx->owner[(size_t)p * C + c] == ctx->shard_rank;
                t.split = (!ctx->owner.empty() && ctx->owner[(size_t)p * C + c] < 0 && ctx->geom_r[K] > 0) ? 1 : 0;
                t.fg_row = (uint32_t)fg_rows; t.fsg_row = (uint32_t)fsg_rows; t.fh_row = (uint32_t)fh_rows; t.t_row = (uint32_t)t_rows;
                // scanned messages: of every forest the programs may push (<= s_prog nodes) and of the children
                // forests the inside recipes read (< s nodes)
                t.nfsg = t.fi.base[std::min(t.fi.s, std::max(t.s_prog, t.fi.s - 1)) + 1];
                if (!t.mine) continue;
                const size_t a = (size_t)M * t.fi.nf, a2 = (size_t)M * t.nfsg, b = (size_t)M * K * t.fi.nfh;
                const size_t tt = ctx->geom_r[K] > 0 ? (size_t)M * ctx->pairs[K].nt : 0;
                fg_rows += a; fsg_rows += a2; fh_rows += b; t_rows += tt;
                per_k[K] += (double)(a + a2 + b + tt) * L * sizeof(float);
            }
        const double bytes = (double)(fg_rows + fsg_rows + fh_rows + t_rows) * L * sizeof(float);
        ctx->table_bytes = bytes;
        if ((bytes <= budget && fg_rows < 0xffffffffull && fh_rows < 0xffffffffull && t_rows < 0xffffffffull) || (env_s && env_sh)) break;
        int worst = 0;                                 // step the geometry of the hungriest cluster count down
        for (int K = 1; K <= KMAX; ++K) if (per_k[K] > per_k[worst]) worst = K;
        bool stepped = false;
        if (worst && ctx->geom_r[worst] > 0) {         // smaller contexts first (more irregular trees), then no pair path
            if (ctx->geom_s[worst] + ctx->geom_r[worst] > worst) --ctx->geom_r[worst];
            else {
                ctx->geom_r[worst] = 0;
                ctx->geom_s[worst] = program_forest_s(worst, ctx->geom_s[worst]);
            }
            stepped = true;
        } else if (worst) stepped = reduce_forest_geometry(&ctx->geom_s[worst], &ctx->geom_sh[worst]);
        if (!stepped) {
            bool any = false;
            for (int K = 1; K <= KMAX && !any; ++K)
                if (k_present[K] && ctx->geom_r[K] == 0) any = reduce_forest_geometry(&ctx->geom_s[K], &ctx->geom_sh[K]);
            if (!any) return ctx->fail(BAMSE_ENOMEM, "bamse_upload_problem: the smallest tables (%.1f GB) exceed the table budget (%.1f GB)", bytes / 1e9, budget / 1e9);
        }
    }
    ctx->slots_needed = 2;
    ctx->smax = 1; ctx->rmax = 0; ctx->max_fh_items = 0; ctx->max_top_rows = 0;
    for (int n = 0; n <= FT_SMAX + 1; ++n) ctx->max_level_rows[n] = 0;
    for (int q = 0; q <= PR_RMAX + 1; ++q) ctx->max_ctx_rows[q] = 0;
    for (int K = 1; K <= KMAX; ++K) {
        if (!k_present[K]) continue;
        const ForestIndex fi = make_forest_index(K, ctx->geom_s[K], ctx->geom_sh[K]);
        const int sp = program_forest_s(K, fi.s);
        ctx->slots_needed = std::max(ctx->slots_needed, forest_slots_needed(K, sp));
        ctx->smax = std::max(ctx->smax, fi.s);
        for (int n = 2; n <= fi.s; ++n) ctx->max_level_rows[n] = std::max(ctx->max_level_rows[n], fi.base[n + 1] - fi.base[n]);
        if (sp == fi.s) ctx->max_top_rows = std::max(ctx->max_top_rows, fi.base[fi.s + 1] - fi.base[fi.s]);
        ctx->max_fh_items = std::max(ctx->max_fh_items, K * fi.nfh);
        if (ctx->geom_r[K] > 0) {
            const bamse_ctx::PairState& ps = ctx->pairs[K];
            ctx->rmax = std::max(ctx->rmax, ctx->geom_r[K]);
            for (int q = 1; q <= ctx->geom_r[K]; ++q)
                ctx->max_ctx_rows[q] = std::max(ctx->max_ctx_rows[q], (int)(ps.lbase[q + 1] - ps.lbase[q]));
        }
    }
    ctx->fh_rows = fh_rows;
    ctx->interp_ready = false;
    BAMSE_CUDA_TRY(ctx, ctx->d_FG.ensure(sizeof(float) * std::max<size_t>(fg_rows, 1) * L));
    BAMSE_CUDA_TRY(ctx, ctx->d_FSG.ensure(sizeof(float) * std::max<size_t>(fsg_rows, 1) * L));
    BAMSE_CUDA_TRY(ctx, ctx->d_T.ensure(sizeof(float) * std::max<size_t>(t_rows, 1) * L));
    BAMSE_CUDA_TRY(ctx, ctx->d_tb.ensure(sizeof(TableBase) * tbs.size()));
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_tb.p, tbs.data(), sizeof(TableBase) * tbs.size(), cudaMemcpyHostToDevice, s));
    {   // the table kernels run over the (p, c, m) triples of this rank only
        std::vector<int32_t> pcm;
        for (int pc = 0; pc < P * C; ++pc)
            if (tbs[(size_t)pc].mine) for (int m = 0; m < M; ++m) pcm.push_back(pc * M + m);
        ctx->n_pcm = (int)pcm.size();
        BAMSE_CUDA_TRY(ctx, ctx->d_pcm.ensure(sizeof(int32_t) * std::max<size_t>(pcm.size(), 1)));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pcm.p, pcm.data(), sizeof(int32_t) * pcm.size(), cudaMemcpyHostToDevice, s));
        BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));              // pcm dies here
    }
    if (!ctx->owner.empty()) {
        BAMSE_CUDA_TRY(ctx, ctx->d_owner.ensure(sizeof(int32_t) * ctx->owner.size()));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_owner.p, ctx->owner.data(), sizeof(int32_t) * ctx->owner.size(), cudaMemcpyHostToDevice, s));
    }
    {
        const ForestRecipe* rptrs[KMAX + 1];
        const uint32_t* optrs[KMAX + 1];
        PairK pks[KMAX + 1];
        memset(pks, 0, sizeof pks);
        for (int K = 0; K <= KMAX; ++K) {
            rptrs[K] = nullptr; optrs[K] = nullptr;
            if (K == 0 || !k_present[K]) continue;
            bamse_ctx::Recipes& r = ctx->recipes[K];
            if (!r.buf.p || r.s != ctx->geom_s[K]) {
                const ForestIndex fiK = make_forest_index(K, ctx->geom_s[K], ctx->geom_sh[K]);
                const std::vector<ForestRecipe> rec = make_forest_recipes(fiK);
                const std::vector<uint32_t> ord = make_level_order(fiK, rec);
                BAMSE_CUDA_TRY(ctx, r.buf.ensure(sizeof(ForestRecipe) * rec.size()));
                BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(r.buf.p, rec.data(), sizeof(ForestRecipe) * rec.size(), cudaMemcpyHostToDevice, s));
                BAMSE_CUDA_TRY(ctx, r.order.ensure(sizeof(uint32_t) * ord.size()));
                BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(r.order.p, ord.data(), sizeof(uint32_t) * ord.size(), cudaMemcpyHostToDevice, s));
                BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));      // rec / ord die here (once per K and geometry)
                r.s = ctx->geom_s[K];
            }
            rptrs[K] = r.buf.as<ForestRecipe>();
            optrs[K] = r.order.as<uint32_t>();
            if (ctx->geom_r[K] > 0) {
                const bamse_ctx::PairState& ps = ctx->pairs[K];
                PairK& pk = pks[K];
                pk.g = make_pair_geom(K, ctx->geom_s[K], ctx->geom_r[K]);
                pk.nt = ps.nt;
                for (int q = 0; q <= PR_RMAX + 1; ++q) pk.lbase[q] = ps.lbase[q];
                pk.map = ps.map.as<uint32_t>();
                pk.recipes = ps.recipes.as<CtxRecipe>();
                pk.pair_tab = ps.has_tab ? ps.tab.as<uint64_t>() : nullptr;
                pk.irr_progs = ps.irr_progs.as<FastProgram>();
                pk.mask_owner = nullptr;
                if (!ctx->owner.empty() && ctx->shard_world > 1 && !ps.mask_trees.empty()) {
                    // node sets to ranks, heaviest first to the least loaded rank (trees per set from the pair table)
                    bamse_ctx::PairState& pw = ctx->pairs[K];
                    if (pw.owner_world != ctx->shard_world) {
                        const int W = ctx->shard_world;
                        const size_t n_masks = pw.mask_trees.size();
                        std::vector<uint32_t> order(n_masks);
                        for (size_t i = 0; i < n_masks; ++i) order[i] = (uint32_t)i;
                        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return pw.mask_trees[a] > pw.mask_trees[b]; });
                        std::vector<double> load(W, 0.0);
                        std::vector<uint8_t> own(n_masks, 0);
                        for (uint32_t m : order) {
                            int r_min = 0;
                            for (int r2 = 1; r2 < W; ++r2) if (load[r2] < load[r_min]) r_min = r2;
                            own[m] = (uint8_t)r_min;
                            load[r_min] += (double)pw.mask_trees[m] + 1.0;
                        }
                        BAMSE_CUDA_TRY(ctx, pw.mask_owner.ensure(n_masks));
                        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(pw.mask_owner.p, own.data(), n_masks, cudaMemcpyHostToDevice, s));
                        BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));
                        pw.owner_world = W;
                    }
                    pk.mask_owner = pw.mask_owner.as<uint8_t>();
                }
                pk.sorted_tree = ps.sorted_tree.as<uint32_t>();
                pk.sorted_entry = ps.sorted_entry.as<uint64_t>();
            }
        }
        BAMSE_CUDA_TRY(ctx, ctx->d_recipe_ptrs.ensure(sizeof rptrs));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_recipe_ptrs.p, rptrs, sizeof rptrs, cudaMemcpyHostToDevice, s));
        BAMSE_CUDA_TRY(ctx, ctx->d_order_ptrs.ensure(sizeof optrs));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_order_ptrs.p, optrs, sizeof optrs, cudaMemcpyHostToDevice, s));
        BAMSE_CUDA_TRY(ctx, ctx->d_pk.ensure(sizeof pks));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pk.p, pks, sizeof pks, cudaMemcpyHostToDevice, s));
        BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));              // the pointer tables die here
    }
    {   // program tables for the small K of this problem (K^(K-1) rows of 64 bytes; K = 8: 134 MB)
        const char* off = getenv("BAMSE_PROGTAB");          // BAMSE_PROGTAB=0: always decode in the kernel
        const bool enable = !(off && atoi(off) == 0);
        const FastProgram* ptrs[KMAX + 1];
        for (int k = 0; k <= KMAX; ++k) ptrs[k] = nullptr;
        ctx->have_progtabs = false;
        for (int c = 0; c < C && enable; ++c) {
            const int K = ctx->K_of_c[c];
            if (K > PROGTAB_KMAX || ptrs[K]) continue;
            // a cluster count whose trees are all regular never runs a program in an enumeration
            if (ctx->geom_r[K] > 0 && ctx->pairs[K].n_irregular == 0) continue;
            bamse_ctx::ProgTab& t = ctx->progtab[K];
            const uint64_t n = ipow_u64(K, K - 1);
            const int sp = program_forest_s(K, ctx->geom_s[K]);
            if (!t.buf.p || t.s != sp || t.sh != ctx->geom_sh[K]) {
                BAMSE_CUDA_TRY(ctx, t.buf.ensure(sizeof(FastProgram) * n));
                launch_build_prog_table(make_forest_index(K, sp, ctx->geom_sh[K]), n, ctx->d_lut.as<int16_t>(),
                                        t.buf.as<FastProgram>(), s);
                ctx->launches++;
                t.s = sp; t.sh = ctx->geom_sh[K];
            }
            ptrs[K] = t.buf.as<FastProgram>();
            ctx->have_progtabs = true;
        }
        BAMSE_CUDA_TRY(ctx, ctx->d_progtabs.ensure(sizeof ptrs));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_progtabs.p, ptrs, sizeof ptrs, cudaMemcpyHostToDevice, s));
        BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    }
    BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));   // the host arrays above (table bases, pointer tables) die at return
    return BAMSE_OK;
}

// Every table of the resident problem from the read counts in device memory: binomial tables, inside tables
// level by level, context tables level by level.  Asynchronous on the ctx stream.  The finishing tables and
// the top-level scanned messages, which only the table-driven programs read, follow on demand (ensure_interp).
static int launch_table_build(bamse_ctx* ctx) {
    cudaStream_t s = ctx->stream;
    unsigned long long* stats = nullptr;
    if (ctx->collect_stats) {
        const bool fresh = ctx->d_scalars.p == nullptr;
        BAMSE_CUDA_TRY(ctx, ctx->d_scalars.ensure(sizeof(uint64_t) * 16));
        if (fresh) BAMSE_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars.p, 0, sizeof(uint64_t) * 16, s));
        stats = reinterpret_cast<unsigned long long*>(ctx->d_scalars.as<uint64_t>() + 4);
    }
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!ctx->ev_free.empty()) { ev = ctx->ev_free.back(); ctx->ev_free.pop_back(); }
    else {
        BAMSE_CUDA_TRY(ctx, cudaEventCreate(&ev.first));
        BAMSE_CUDA_TRY(ctx, cudaEventCreate(&ev.second));
    }
    BAMSE_CUDA_TRY(ctx, cudaEventRecord(ev.first, s));
    launch_tables(ctx->view(), ctx->d_var.as<double>(), ctx->d_ref.as<double>(), ctx->d_logc.as<double>(),
                  ctx->d_logcf.as<double>(), ctx->d_log1mcf.as<double>(), ctx->d_D64.as<double>(),
                  ctx->d_D32.as<float>(), s);
    launch_tables_fast(ctx->view(), ctx->fast(), ctx->d_D2.as<float>(), ctx->d_FG.as<float>(), ctx->d_FSG.as<float>(),
                       ctx->d_SD2.as<float>(), s);
    ctx->launches += 2;
    for (int n = 2; n <= ctx->smax; ++n) {
        launch_forest_level(ctx->view(), ctx->fast(), n, ctx->max_level_rows[n], ctx->d_FG.as<float>(), ctx->d_FSG.as<float>(), 0, stats, s);
        ctx->launches++;
    }
    for (int q = 1; q <= ctx->rmax; ++q) {
        if (ctx->max_ctx_rows[q] <= 0) continue;
        launch_ctx_level(ctx->view(), ctx->fast(), q, ctx->max_ctx_rows[q], ctx->d_T.as<float>(), stats, s);
        ctx->launches++;
    }
    BAMSE_CUDA_TRY(ctx, cudaEventRecord(ev.second, s));
    ctx->ev_build_used.push_back(ev);
    if (ctx->ev_build_used.size() > 256) {           // nobody is reading the timer: keep the newest only
        ctx->ev_free.push_back(ctx->ev_build_used.front());
        ctx->ev_build_used.erase(ctx->ev_build_used.begin());
    }
    ctx->interp_ready = false;
    BAMSE_CUDA_TRY(ctx, cudaGetLastError());
    return BAMSE_OK;
}

// What only the table-driven programs read: the scanned messages of the top-level forests and the finishing tables.
static int ensure_interp(bamse_ctx* ctx) {
    if (ctx->interp_ready) return BAMSE_OK;
    cudaStream_t s = ctx->stream;
    BAMSE_CUDA_TRY(ctx, ctx->d_FH.ensure(sizeof(float) * std::max<size_t>(ctx->fh_rows, 1) * L));
    if (ctx->max_top_rows > 0) {
        launch_forest_scan_top(ctx->view(), ctx->fast(), ctx->max_top_rows, ctx->d_FG.as<float>(), ctx->d_FSG.as<float>(), s);
        ctx->launches++;
    }
    launch_forest_fh(ctx->view(), ctx->fast(), ctx->max_fh_items, ctx->d_FH.as<float>(), s);
    ctx->launches++;
    BAMSE_CUDA_TRY(ctx, cudaGetLastError());
    ctx->interp_ready = true;
    return BAMSE_OK;
}

extern "C" {

int bamse_abi_version(void) { return BAMSE_ABI_VERSION; }

const char* bamse_last_error(const bamse_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int bamse_create(bamse_ctx** out, int device_ordinal) {
    if (!out) { g_create_error = "bamse_create: out is NULL"; return BAMSE_EINVAL; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("bamse_create: no CUDA device (") + cudaGetErrorString(e) +
                         "); libbamse_cuda has no CPU fallback";
        cudaGetLastError();
        return BAMSE_ENODEV;
    }
    if (device_ordinal < 0 || device_ordinal >= n) {
        g_create_error = "bamse_create: device ordinal out of range";
        return BAMSE_EINVAL;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_ordinal)) != cudaSuccess) {
        g_create_error = std::string("bamse_create: ") + cudaGetErrorString(e);
        return BAMSE_ECUDA;
    }
    if (prop.major != 10) {
        char b[160];
        snprintf(b, sizeof b, "bamse_create: device %d is sm_%d%d; this library is built for sm_100a only",
                 device_ordinal, prop.major, prop.minor);
        g_create_error = b;
        return BAMSE_ENODEV;
    }
    bamse_ctx* c = new (std::nothrow) bamse_ctx();
    if (!c) { g_create_error = "bamse_create: out of host memory"; return BAMSE_ENOMEM; }
    c->device = device_ordinal;
    cudaSetDevice(device_ordinal);
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("bamse_create: ") + cudaGetErrorString(e);
        delete c;
        return BAMSE_ECUDA;
    }
    c->stream = c->own_stream;
    *out = c;
    return BAMSE_OK;
}

void bamse_destroy(bamse_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DBuf* bufs[] = {&ctx->d_K, &ctx->d_p0, &ctx->d_var, &ctx->d_ref, &ctx->d_logc, &ctx->d_logcf, &ctx->d_log1mcf,
                    &ctx->d_D64, &ctx->d_D32, &ctx->d_D2, &ctx->d_SD2, &ctx->d_FG, &ctx->d_FSG, &ctx->d_FH, &ctx->d_tb, &ctx->d_lut,
                    &ctx->d_recipe_ptrs, &ctx->d_segs, &ctx->d_scalars, &ctx->d_cand, &ctx->d_items,
                    &ctx->d_scores, &ctx->d_const, &ctx->d_thr, &ctx->d_begin, &ctx->d_end, &ctx->d_records,
                    &ctx->d_dense0, &ctx->d_dense1, &ctx->d_split, &ctx->d_lists, &ctx->d_T, &ctx->d_pk, &ctx->d_flat,
                    &ctx->d_flat_meta, &ctx->d_lists2, &ctx->d_segs2, &ctx->d_owner, &ctx->d_part, &ctx->d_pcm};
    for (DBuf* b : bufs) b->release();
    for (auto& t : ctx->progtab) t.buf.release();
    for (auto& t : ctx->pairs) { t.tab.release(); t.map.release(); t.recipes.release(); t.sorted_tree.release(); t.sorted_entry.release(); t.irr_progs.release(); t.mask_owner.release(); }
    for (auto& r : ctx->recipes) { r.buf.release(); r.order.release(); }
    ctx->d_order_ptrs.release();
    ctx->d_progtabs.release();
    for (auto* v : {&ctx->ev_used, &ctx->ev_build_used, &ctx->ev_free})
        for (auto& ev : *v) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int bamse_set_stream(bamse_ctx* ctx, void* cuda_stream) {
    CTX_CHECK(ctx);
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return BAMSE_OK;
}

int bamse_set_shard(bamse_ctx* ctx, int rank, int world) {
    CTX_CHECK(ctx);
    if (world < 1 || rank < 0 || rank >= world) return ctx->fail(BAMSE_EINVAL, "bamse_set_shard: need 0 <= rank < world");
    ctx->shard_rank = rank; ctx->shard_world = world;
    // tables exist only for the (parameter set, clustering) pairs the plan gave this rank: a resident problem
    // that was planned for another shard has to be uploaded again
    if (ctx->have_problem && !ctx->owner.empty() && (rank != ctx->plan_rank || world != ctx->plan_world)) ctx->have_problem = false;
    return BAMSE_OK;
}

int bamse_fast_variant(const bamse_ctx* ctx) { return ctx ? ctx->use_framed : -1; }

int bamse_debug_shard_plan(int P, int C, const int32_t* K_of_c, int world, const double* build_cost, const double* enum_cost,
                           int32_t* out_owner) try {
    if (P < 1 || C < 1 || world < 1 || !K_of_c || !build_cost || !enum_cost || !out_owner) return BAMSE_EINVAL;
    for (int c = 0; c < C; ++c) if (K_of_c[c] < 1 || K_of_c[c] > KMAX) return BAMSE_EINVAL;
    const std::vector<int32_t> owner = plan_owners(P, C, K_of_c, world, build_cost, enum_cost);
    for (size_t i = 0; i < owner.size(); ++i) out_owner[i] = owner[i];
    return BAMSE_OK;
} catch (...) { return on_exception(nullptr, "bamse_debug_shard_plan"); }

int bamse_debug_pair_stats(bamse_ctx* ctx, int K, int s, int r, int64_t* out4) try {
    CTX_CHECK(ctx);
    if (K < 2 || K > KMAX || s < 1 || s > FT_SMAX || r < 1 || r > PR_RMAX || r - 1 > s || s + r < K || !out4)
        return ctx->fail(BAMSE_EINVAL, "bamse_debug_pair_stats: bad geometry");
    if (!ctx->d_lut.p) {
        BAMSE_CUDA_TRY(ctx, ctx->d_lut.ensure(sizeof(int16_t) * FT_LUT_SIZE));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_lut.p, forest_lut().lut.data(), sizeof(int16_t) * FT_LUT_SIZE, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->have_problem = false;                       // the resident problem may have been planned with another geometry
    ctx->plan_valid = false;
    {
        int s0 = 1, sh0 = 1, r0 = 0;
        default_forest_geometry(K, &s0, &sh0, &r0);
        ctx->geom_sh[K] = std::min(sh0, program_forest_s(K, s));
    }
    ctx->geom_s[K] = s; ctx->geom_r[K] = r;
    const int rc = ensure_pair_state(ctx, K);
    if (rc) return rc;
    out4[0] = make_forest_index(K, s, 1).nf;
    out4[1] = make_pair_geom(K, s, r).nt;
    out4[2] = ctx->pairs[K].nt;
    out4[3] = (int64_t)ctx->pairs[K].n_irregular;
    return BAMSE_OK;
} catch (...) { return on_exception(ctx, "bamse_debug_pair_stats"); }

int bamse_get_shard_plan(bamse_ctx* ctx, int32_t* out_owner) {
    CTX_CHECK(ctx);
    if (!ctx->have_problem) return ctx->fail(BAMSE_ESTATE, "bamse_get_shard_plan: no problem uploaded");
    if (!out_owner) return ctx->fail(BAMSE_EINVAL, "bamse_get_shard_plan: out_owner is NULL");
    for (size_t i = 0; i < (size_t)ctx->P * ctx->C; ++i) out_owner[i] = ctx->owner.empty() ? -1 : ctx->owner[i];
    return BAMSE_OK;
}

int bamse_set_option(bamse_ctx* ctx, int option, int value) {
    CTX_CHECK(ctx);
    if (option == BAMSE_OPT_EARLY_EXIT) { ctx->early_exit = value ? 1 : 0; return BAMSE_OK; }
    if (option == BAMSE_OPT_SHARD_PLAN) { ctx->shard_plan = value ? 1 : 0; return BAMSE_OK; }
    if (option == BAMSE_OPT_STATS) { ctx->collect_stats = value ? 1 : 0; return BAMSE_OK; }
    return ctx->fail(BAMSE_EINVAL, "bamse_set_option: unknown option %d", option);
}

int64_t bamse_samples_evaluated(bamse_ctx* ctx, int reset) {
    if (!ctx || !ctx->d_scalars.p) return 0;
    cudaSetDevice(ctx->device);
    uint64_t v = 0;
    uint64_t* scal = ctx->d_scalars.as<uint64_t>();
    if (cudaMemcpyAsync(&v, scal + 8, sizeof v, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (reset) cudaMemsetAsync(scal + 8, 0, sizeof v, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return (int64_t)v;
}

int64_t bamse_launch_count(const bamse_ctx* ctx) { return ctx ? ctx->launches : 0; }
int bamse_last_topk_verified(const bamse_ctx* ctx) { return ctx ? ctx->last_verified : 0; }
int bamse_last_topk_preselected(const bamse_ctx* ctx) { return ctx ? ctx->last_ksel : 0; }
int64_t bamse_last_eval_count(const bamse_ctx* ctx) { return ctx ? ctx->last_evals : 0; }

int bamse_kernel_time(bamse_ctx* ctx, int reset, double* out_ms, int64_t* out_launches) try {
    CTX_CHECK(ctx);
    BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0.0;
    for (auto& ev : ctx->ev_used) {
        float t = 0.f;
        BAMSE_CUDA_TRY(ctx, cudaEventElapsedTime(&t, ev.first, ev.second));
        ms += t;
    }
    if (out_ms) *out_ms = ms;
    if (out_launches) *out_launches = (int64_t)ctx->ev_used.size();
    if (reset) {
        for (auto& ev : ctx->ev_used) ctx->ev_free.push_back(ev);
        ctx->ev_used.clear();
    }
    return BAMSE_OK;
} catch (...) { return on_exception(ctx, "bamse_kernel_time"); }

int bamse_table_time(bamse_ctx* ctx, int reset, double* out_ms, int64_t* out_builds) try {
    CTX_CHECK(ctx);
    BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0.0;
    for (auto& ev : ctx->ev_build_used) {
        float t = 0.f;
        BAMSE_CUDA_TRY(ctx, cudaEventElapsedTime(&t, ev.first, ev.second));
        ms += t;
    }
    if (out_ms) *out_ms = ms;
    if (out_builds) *out_builds = (int64_t)ctx->ev_build_used.size();
    if (reset) {
        for (auto& ev : ctx->ev_build_used) ctx->ev_free.push_back(ev);
        ctx->ev_build_used.clear();
    }
    return BAMSE_OK;
} catch (...) { return on_exception(ctx, "bamse_table_time"); }

int bamse_kernel_stats(bamse_ctx* ctx, int reset, uint64_t* out4) {
    CTX_CHECK(ctx);
    if (!out4) return ctx->fail(BAMSE_EINVAL, "bamse_kernel_stats: out4 is NULL");
    for (int i = 0; i < 4; ++i) out4[i] = 0;
    if (!ctx->d_scalars.p) return BAMSE_OK;
    uint64_t* scal = ctx->d_scalars.as<uint64_t>();
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(out4, scal + 4, sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) BAMSE_CUDA_TRY(ctx, cudaMemsetAsync(scal + 4, 0, sizeof(uint64_t) * 4, ctx->stream));
    BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BAMSE_OK;
}

int bamse_decode_tree(int K, uint64_t tree_index, int8_t* out_parent) {
    if (K < 1 || K > KMAX || !out_parent) return BAMSE_EINVAL;
    if (tree_index >= ipow_u64(K, K - 1)) return BAMSE_EINVAL;
    decode_tree(K, tree_index, out_parent);
    return BAMSE_OK;
}

// ---- introspection of the data-independent machinery (tests run these without a GPU) ----------------------
int bamse_debug_forest_geometry(int K, int32_t* out_s, int32_t* out_sh, int32_t* out_nf, int32_t* out_nfh, int32_t* out_slots) {
    if (K < 1 || K > KMAX) return BAMSE_EINVAL;
    int s = 1, sh = 1;
    if (out_s && *out_s > 0 && out_sh && *out_sh > 0) { s = *out_s; sh = *out_sh; }     // caller-given geometry
    else { default_forest_geometry(K, &s, &sh); s = program_forest_s(K, s); if (sh > s) sh = s; }   // what the programs see
    if (s > FT_SMAX || sh > s) return BAMSE_EINVAL;
    const ForestIndex fi = make_forest_index(K, s, sh);
    if (out_s) *out_s = s;
    if (out_sh) *out_sh = sh;
    if (out_nf) *out_nf = fi.nf;
    if (out_nfh) *out_nfh = fi.nfh;
    if (out_slots) *out_slots = forest_slots_needed(K, s);
    return BAMSE_OK;
}

int bamse_debug_forest_recipes(int K, int s, int32_t* out /* [nf][5]: kind, u, a, b, mask */) try {
    if (K < 1 || K > KMAX || s < 1 || s > FT_SMAX || !out) return BAMSE_EINVAL;
    const std::vector<ForestRecipe> rec = make_forest_recipes(make_forest_index(K, s, 1));
    for (size_t i = 0; i < rec.size(); ++i) {
        out[5 * i] = rec[i].kind; out[5 * i + 1] = rec[i].u; out[5 * i + 2] = rec[i].a; out[5 * i + 3] = rec[i].b;
        out[5 * i + 4] = rec[i].mask;
    }
    return BAMSE_OK;
} catch (...) { return on_exception(nullptr, "bamse_debug_forest_recipes"); }

int bamse_debug_program(int k, const int8_t* parent, const int8_t* nodes, int Kc, int s, int sh,
                        int32_t* out /* [6 + 2 * 20]: n_ops, fin, fin_arg, fin_r, scre, max_sp, ops[20], args[20] */) try {
    if (k < 1 || k > KMAX || Kc < k || Kc > KMAX || s < 1 || s > FT_SMAX || sh < 1 || sh > s || !parent || !out) return BAMSE_EINVAL;
    FastProgram prog;
    memset(&prog, 0, sizeof prog);
    if (!build_program_forest(k, parent, nodes, make_forest_index(Kc, s, sh), forest_lut().lut.data(), &prog)) return BAMSE_EINVAL;
    out[0] = prog.n_ops; out[1] = prog.fin; out[2] = prog.fin_arg; out[3] = prog.fin_r; out[4] = prog.scre; out[5] = prog.max_sp;
    for (int i = 0; i < FAST_MAX_OPS; ++i) {
        out[6 + i] = i < prog.n_ops ? fast_op(prog, i) : -1;
        out[6 + FAST_MAX_OPS + i] = i < prog.n_ops ? prog.arg[i] : 0;
    }
    return BAMSE_OK;
} catch (...) { return on_exception(nullptr, "bamse_debug_program"); }

int bamse_debug_pair_geometry(int K, int32_t* s, int32_t* r, int32_t* out_nf, int32_t* out_nt) {
    if (K < 1 || K > KMAX || !s || !r) return BAMSE_EINVAL;
    if (*s <= 0 || *r < 0) { int sh = 1, rr = 0, ss = 1; default_forest_geometry(K, &ss, &sh, &rr); *s = ss; *r = rr; }
    if (*s > FT_SMAX || *r > PR_RMAX || (*r > 0 && *r - 1 > *s)) return BAMSE_EINVAL;
    if (out_nf) *out_nf = make_forest_index(K, *s, 1).nf;
    if (out_nt) *out_nt = (int32_t)make_pair_geom(K, *s, *r).nt;
    return BAMSE_OK;
}

int bamse_debug_ctx_recipes(int K, int s, int r, int32_t* out /* [nt][3]: parent context row, FG row of H, marked label */) try {
    if (K < 1 || K > KMAX || s < 1 || s > FT_SMAX || r < 1 || r > PR_RMAX || r - 1 > s || !out) return BAMSE_EINVAL;
    const std::vector<CtxRecipe> rec = make_ctx_recipes(make_forest_index(K, s, 1), make_pair_geom(K, s, r));
    for (size_t i = 0; i < rec.size(); ++i) { out[3 * i] = rec[i].parent_row; out[3 * i + 1] = rec[i].h_row; out[3 * i + 2] = rec[i].v; }
    return BAMSE_OK;
} catch (...) { return on_exception(nullptr, "bamse_debug_ctx_recipes"); }

int bamse_debug_classify(int K, const int8_t* parent, int s, int r, int32_t* out /* regular, FG row, context row */) try {
    if (K < 1 || K > KMAX || s < 1 || s > FT_SMAX || r < 0 || r > PR_RMAX || !parent || !out) return BAMSE_EINVAL;
    TreeProgram ref;
    if (!build_program(K, parent, nullptr, &ref)) return BAMSE_EINVAL;
    uint32_t fg = 0, t = 0;
    const bool reg = r > 0 && classify_tree(K, parent, make_forest_index(K, s, 1), forest_lut().lut.data(), make_pair_geom(K, s, r), &fg, &t);
    out[0] = reg ? 1 : 0; out[1] = (int32_t)fg; out[2] = (int32_t)t;
    return BAMSE_OK;
} catch (...) { return on_exception(nullptr, "bamse_debug_classify"); }

int bamse_upload_problem(bamse_ctx* ctx, int P, const double* err, int C, int M, int Kmax, const int32_t* K_of_c,
                         const int64_t* var_counts, const int64_t* ref_counts) try {
    CTX_CHECK(ctx);
    ctx->have_problem = false;
    if (P < 1 || C < 1 || M < 1 || M > BAMSE_MMAX || Kmax < 1 || Kmax > KMAX || !err || !K_of_c || !var_counts ||
        !ref_counts)
        return ctx->fail(BAMSE_EINVAL, "bamse_upload_problem: bad sizes (P=%d C=%d M=%d Kmax=%d) or NULL pointer", P,
                         C, M, Kmax);
    for (int p = 0; p < P; ++p)
        if (!(err[p] > 0.0) || !(err[p] < 0.5))
            return ctx->fail(BAMSE_EINVAL, "bamse_upload_problem: err[%d]=%g must be in (0, 0.5)", p, err[p]);
    for (int c = 0; c < C; ++c)
        if (K_of_c[c] < 1 || K_of_c[c] > Kmax)
            return ctx->fail(BAMSE_EINVAL, "bamse_upload_problem: K_of_c[%d]=%d out of [1,%d]", c, K_of_c[c], Kmax);
    const size_t ncnt = (size_t)C * Kmax * M;
    std::vector<double> var(ncnt, 0.0), ref(ncnt, 0.0), logc(ncnt, 0.0), p0(C);
    for (int c = 0; c < C; ++c) {
        const int K = K_of_c[c];
        p0[c] = 1.0 / K * (0.0 - lgamma((double)K));                         // clustering.py:29
        for (int k = 0; k < K; ++k)
            for (int m = 0; m < M; ++m) {
                const size_t i = ((size_t)c * Kmax + k) * M + m;
                if (var_counts[i] < 0 || ref_counts[i] < 0)
                    return ctx->fail(BAMSE_EINVAL, "bamse_upload_problem: negative read count at c=%d k=%d m=%d", c,
                                     k, m);
            }
    }
    // the host's share of distrib (clustering.py:21-23): the three lgamma terms per count pair and the logs of the
    // grid, in fp64 with the C library's functions; a few worker threads when there are tens of thousands
    std::vector<double> logcf((size_t)P * L), log1mcf((size_t)P * L);
    {
        const double step = 1.0 / (double)(L - 1);
        auto counts_part = [&](int c0, int c1) {
            for (int c = c0; c < c1; ++c)
                for (int k = 0; k < K_of_c[c]; ++k)
                    for (int m = 0; m < M; ++m) {
                        const size_t i = ((size_t)c * Kmax + k) * M + m;
                        const int64_t a = var_counts[i], b = ref_counts[i];
                        var[i] = (double)a; ref[i] = (double)b;
                        // gammaln(a+b+1) - gammaln(a+1) - gammaln(b+1)          clustering.py:23
                        logc[i] = lgamma((double)(a + b) + 1.0) - lgamma((double)a + 1.0) - lgamma((double)b + 1.0);
                    }
        };
        auto grid_part = [&](int p0_, int p1_) {
            for (int p = p0_; p < p1_; ++p)
                for (int x = 0; x < L; ++x) {
                    const double gx = (x == L - 1) ? 1.0 : (double)x * step;      // numpy linspace
                    volatile double scaled = gx * (0.5 - err[p]);                  // no FMA contraction
                    const double cf = scaled + err[p];
                    logcf[(size_t)p * L + x] = log(cf);
                    log1mcf[(size_t)p * L + x] = log(1.0 - cf);
                }
        };
        const size_t work = ncnt * 3 + (size_t)P * L * 2;
        const int n_thr = work < 20000 ? 1 : (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        if (n_thr == 1) { counts_part(0, C); grid_part(0, P); }
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < n_thr; ++t)
                pool.emplace_back([&, t]() {
                    counts_part((int)((int64_t)C * t / n_thr), (int)((int64_t)C * (t + 1) / n_thr));
                    grid_part((int)((int64_t)P * t / n_thr), (int)((int64_t)P * (t + 1) / n_thr));
                });
            for (auto& th : pool) th.join();
        }
    }
    const size_t nrows = (size_t)P * C * M * Kmax;
    BAMSE_CUDA_TRY(ctx, ctx->d_K.ensure(sizeof(int32_t) * C));
    BAMSE_CUDA_TRY(ctx, ctx->d_p0.ensure(sizeof(double) * C));
    BAMSE_CUDA_TRY(ctx, ctx->d_var.ensure(sizeof(double) * ncnt));
    BAMSE_CUDA_TRY(ctx, ctx->d_ref.ensure(sizeof(double) * ncnt));
    BAMSE_CUDA_TRY(ctx, ctx->d_logc.ensure(sizeof(double) * ncnt));
    BAMSE_CUDA_TRY(ctx, ctx->d_logcf.ensure(sizeof(double) * P * L));
    BAMSE_CUDA_TRY(ctx, ctx->d_log1mcf.ensure(sizeof(double) * P * L));
    BAMSE_CUDA_TRY(ctx, ctx->d_D64.ensure(sizeof(double) * nrows * L));
    BAMSE_CUDA_TRY(ctx, ctx->d_D32.ensure(sizeof(float) * nrows * L));
    BAMSE_CUDA_TRY(ctx, ctx->d_D2.ensure(sizeof(float) * nrows * L));
    BAMSE_CUDA_TRY(ctx, ctx->d_SD2.ensure(sizeof(float) * nrows * L));
    ctx->P = P; ctx->C = C; ctx->M = M; ctx->Kmax = Kmax;
    ctx->K_of_c.assign(K_of_c, K_of_c + C);
    for (int K = 0; K <= KMAX; ++K) ctx->k_present[K] = false;
    for (int c = 0; c < C; ++c) ctx->k_present[K_of_c[c]] = true;
    cudaStream_t s = ctx->stream;
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_K.p, K_of_c, sizeof(int32_t) * C, cudaMemcpyHostToDevice, s));
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_p0.p, p0.data(), sizeof(double) * C, cudaMemcpyHostToDevice, s));
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_var.p, var.data(), sizeof(double) * ncnt, cudaMemcpyHostToDevice, s));
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_ref.p, ref.data(), sizeof(double) * ncnt, cudaMemcpyHostToDevice, s));
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_logc.p, logc.data(), sizeof(double) * ncnt, cudaMemcpyHostToDevice, s));
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_logcf.p, logcf.data(), sizeof(double) * P * L, cudaMemcpyHostToDevice, s));
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_log1mcf.p, log1mcf.data(), sizeof(double) * P * L, cudaMemcpyHostToDevice, s));
    // the plan (geometry, table bases, sharding, data-independent tables) depends on the SHAPE of the problem only:
    // a same-shaped problem (other counts / error rates: sweeps, bootstraps, the next batch) reuses it
    std::vector<int32_t> shape = {P, C, M, Kmax, ctx->shard_rank, ctx->shard_world, ctx->shard_plan, pairs_enabled() ? 1 : 0};
    shape.insert(shape.end(), K_of_c, K_of_c + C);
    {   // ... and on the environment switches plan_tables reads
        uint32_t h = 2166136261u;
        for (const char* name : {"BAMSE_FOREST_S", "BAMSE_FOREST_SH", "BAMSE_PAIR_R", "BAMSE_TABLE_GB", "BAMSE_PROGTAB", "BAMSE_SHARD_PLAN"}) {
            const char* v = getenv(name);
            for (const char* q = v ? v : "-"; *q; ++q) { h ^= (unsigned char)*q; h *= 16777619u; }
            h ^= 0xffu; h *= 16777619u;
        }
        shape.push_back((int32_t)h);
    }
    int rc = BAMSE_OK;
    if (!ctx->plan_valid || shape != ctx->plan_shape) {
        ctx->plan_valid = false;
        rc = plan_tables(ctx);
        if (rc) return rc;
        ctx->plan_shape = shape;
        ctx->plan_valid = true;
    }
    rc = launch_table_build(ctx);
    if (rc) return rc;
    BAMSE_CUDA_TRY(ctx, cudaGetLastError());
    BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));   // host staging vectors die at return
    ctx->have_problem = true;
    {   // FNV-1a over the inputs: re-uploading the same problem reuses the probed kernel variant
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](const void* ptr, size_t n) {
            const unsigned char* b = static_cast<const unsigned char*>(ptr);
            for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        };
        mix(err, sizeof(double) * P); mix(K_of_c, sizeof(int32_t) * C);
        mix(var_counts, sizeof(int64_t) * ncnt); mix(ref_counts, sizeof(int64_t) * ncnt);
        ctx->problem_sig = h ^ ((uint64_t)M << 48) ^ ((uint64_t)Kmax << 56);
    }
    ctx->use_framed = (ctx->probed_choice >= 0 && ctx->probed_sig == ctx->problem_sig) ? ctx->probed_choice : -1;
    return BAMSE_OK;
} catch (...) { return on_exception(ctx, "bamse_upload_problem"); }

int bamse_rebuild_tables(bamse_ctx* ctx) try {
    CTX_CHECK(ctx);
    if (!ctx->have_problem) return ctx->fail(BAMSE_ESTATE, "bamse_rebuild_tables: no problem uploaded");
    const int rc = launch_table_build(ctx);
    if (rc) return rc;
    BAMSE_CUDA_TRY(ctx, cudaGetLastError());
    return BAMSE_OK;
} catch (...) { return on_exception(ctx, "bamse_rebuild_tables"); }

// ---- internal: score a host vector of items, results to host ----------------------
static int score_items_host(bamse_ctx* ctx, const std::vector<ScoreItem>& items, int precision, double* out) {
    const size_t n = items.size();
    if (n == 0) return BAMSE_OK;
    // the log-domain kernels return one value per (item, sample); the fp32 production kernel one per item
    const size_t per = precision == BAMSE_FP32 ? 1 : (size_t)ctx->M;
    BAMSE_CUDA_TRY(ctx, ctx->d_items.ensure(sizeof(ScoreItem) * n));
    BAMSE_CUDA_TRY(ctx, ctx->d_scores.ensure(sizeof(double) * n * per));
    cudaStream_t s = ctx->stream;
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_items.p, items.data(), sizeof(ScoreItem) * n, cudaMemcpyHostToDevice, s));
    if (precision == BAMSE_FP32) {
        const int rc = ensure_interp(ctx);
        if (rc) return rc;
        launch_score_items_fast(ctx->view(), ctx->fast(), ctx->d_items.as<ScoreItem>(), (int64_t)n, ctx->d_scores.as<double>(), s);
    }
    else
        launch_score_items(precision, ctx->view(), ctx->d_items.as<ScoreItem>(), (int64_t)n, ctx->d_scores.as<double>(), s);
    ctx->launches++;
    BAMSE_CUDA_TRY(ctx, cudaGetLastError());
    if (per == 1) {
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->d_scores.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
        BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    } else {
        std::vector<double> part(n * per);
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(part.data(), ctx->d_scores.p, sizeof(double) * n * per, cudaMemcpyDeviceToHost, s));
        BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));
        for (size_t i = 0; i < n; ++i) {              // np.sum over samples (clustering.py:89), in sample order
            double t = 0.0;
            for (size_t m = 0; m < per; ++m) t += part[i * per + m];
            out[i] = t;
        }
    }
    return BAMSE_OK;
}

int bamse_score_list(bamse_ctx* ctx, int64_t n_items, const int32_t* item_clust, const int8_t* item_parent,
                     const int8_t* item_nodes, int precision, double* out_score) try {
    CTX_CHECK(ctx);
    if (!ctx->have_problem) return ctx->fail(BAMSE_ESTATE, "bamse_score_list: no problem uploaded");
    if (n_items < 0 || (n_items > 0 && (!item_clust || !item_parent || !item_nodes || !out_score)))
        return ctx->fail(BAMSE_EINVAL, "bamse_score_list: NULL pointer or negative n_items");
    if (!valid_precision(precision))
        return ctx->fail(BAMSE_EINVAL, "bamse_score_list: precision must be BAMSE_FP32 or BAMSE_FP64");
    const int Kmax = ctx->Kmax;
    std::vector<ScoreItem> items((size_t)ctx->P * n_items);
    for (int64_t i = 0; i < n_items; ++i) {
        ScoreItem it;
        memset(&it, 0, sizeof it);
        it.c = item_clust[i];
        if (it.c < 0 || it.c >= ctx->C) return ctx->fail(BAMSE_EINVAL, "bamse_score_list: item %lld: clustering %d out of range", (long long)i, it.c);
        const int Kc = ctx->K_of_c[it.c];
        int k = 0;
        while (k < Kmax && item_nodes[i * Kmax + k] >= 0) ++k;
        if (k < 1) return ctx->fail(BAMSE_EINVAL, "bamse_score_list: item %lld has no nodes", (long long)i);
        uint32_t seen = 0;
        for (int v = 0; v < k; ++v) {
            const int nd = item_nodes[i * Kmax + v];
            if (nd >= Kc || (seen >> nd) & 1u)
                return ctx->fail(BAMSE_EINVAL, "bamse_score_list: item %lld: node id %d invalid or repeated", (long long)i, nd);
            seen |= 1u << nd;
            it.nodes[v] = (int8_t)nd;
            it.parent[v] = item_parent[i * Kmax + v];
        }
        it.k = (int8_t)k;
        TreeProgram prog;
        if (!build_program(k, it.parent, it.nodes, &prog))
            return ctx->fail(BAMSE_EINVAL, "bamse_score_list: item %lld is not a rooted tree (one -1, no cycles)", (long long)i);
        for (int p = 0; p < ctx->P; ++p) { it.p = p; items[(size_t)p * n_items + i] = it; }
    }
    return score_items_host(ctx, items, precision, out_score);
} catch (...) { return on_exception(ctx, "bamse_score_list"); }

static int enumerate_dev(bamse_ctx* ctx, const double* d_const, const double* d_thr, const uint64_t* d_begin,
                         const uint64_t* d_end, int topk, int precision, double slack_abs, double slack_rel,
                         bamse_record* d_out, double* d_dense_score, double* d_dense_logscre, const Segment* host_seg);

// The fp32 kernel has two instantiations: the exponential path only (lean: 10 or 7 CTAs/SM), and one that
// also carries the linear-domain "framed" path (6 CTAs/SM), which wins when the tables are flat
// (low coverage: most terms of a convolution matter) and loses on sharply peaked tables.  Which one
// is faster is a property of the data, so it is measured once per problem on an interleaved sample of
// 8 192 - 40 000 trees (a few ms; only for problems of >= 50 000 trees) and remembered.
// BAMSE_FRAMED=0/1 in the environment overrides.
static int probe_fast_variant(bamse_ctx* ctx) {
    if (const char* f = getenv("BAMSE_FRAMED")) { ctx->use_framed = atoi(f) ? 1 : 0; return BAMSE_OK; }
    const int P = ctx->P, C = ctx->C;
    uint64_t total = 0;
    for (int c = 0; c < C; ++c) total += ipow_u64(ctx->K_of_c[c], ctx->K_of_c[c] - 1);
    total *= (uint64_t)P;
    // small problems: a probe would cost more than it can save -- exponential instantiation
    if (total < 50000) { ctx->use_framed = 0; return BAMSE_OK; }
    // sample: every W-th tree index of every clustering (interleaved like a shard, so every tree shape and
    // every root takes part: the low indices alone are the star-like trees around node 0), total/8 but
    // 8 192 .. 40 000 trees in all -- enough to fill every warp slot of the GPU several times (a shorter run
    // measures latency, not throughput) while both timed runs together cost at most a quarter of one full
    // enumeration.  W is kept coprime to every K <= 12: index = root + K * code, so a stride that shares a
    // factor with K would visit only some of the roots.
    const uint64_t sample = std::min<uint64_t>(40000, std::max<uint64_t>(8192, total / 8));
    uint64_t W = std::max<uint64_t>(1, total / sample);
    while (W > 1 && (W % 2 == 0 || W % 3 == 0 || W % 5 == 0 || W % 7 == 0 || W % 11 == 0)) --W;
    if (W > 0x7fffffffull) W = 0x7fffffffull;        // Segment::stride is 32-bit (2^31 - 1 is prime)
    const int save_rank = ctx->shard_rank, save_world = ctx->shard_world;
    ctx->shard_rank = 0; ctx->shard_world = (int)W;
    ctx->probing = true;
    float best_ms = 0.f;
    int best = 0, rc = BAMSE_OK;
    for (int v = 0; v < 2 && rc == BAMSE_OK; ++v) {
        ctx->use_framed = v;
        rc = enumerate_dev(ctx, nullptr, nullptr, nullptr, nullptr, 0, BAMSE_FP32, 0.0, 0.0, nullptr, nullptr, nullptr,
                           nullptr);
        if (rc) break;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = ctx->fail(BAMSE_ECUDA, "probe: stream sync failed"); break; }
        auto ev = ctx->ev_used.back();
        ctx->ev_used.pop_back();
        ctx->ev_free.push_back(ev);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev.first, ev.second);
        if (v == 0 || ms < best_ms) { best_ms = ms; best = v; }
    }
    ctx->probing = false;
    ctx->shard_rank = save_rank; ctx->shard_world = save_world;
    if (rc) { ctx->use_framed = 0; return rc; }
    ctx->use_framed = best;
    ctx->probed_choice = best;
    ctx->probed_sig = ctx->problem_sig;
    // probe work must not appear in the executed-op statistics of the caller
    BAMSE_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars.as<uint64_t>() + 4, 0, sizeof(uint64_t) * 4, ctx->stream));
    return BAMSE_OK;
}

// ---- internal: run an enumeration; records land in d_out [P][topk] ----------------
// fp32 with the pair path: k_enumerate_pairs over the cluster counts that have context tables (regular trees
// scored, irregular ones listed), k_enumerate_fast in flat mode over that list, k_enumerate_fast in range mode over
// the cluster counts without a pair path; all of them append to one candidate bag, k_merge_topk reduces it.
static int enumerate_dev(bamse_ctx* ctx, const double* d_const, const double* d_thr, const uint64_t* d_begin,
                         const uint64_t* d_end, int topk, int precision, double slack_abs, double slack_rel,
                         bamse_record* d_out, double* d_dense_score, double* d_dense_logscre,
                         const Segment* host_seg /* non-NULL: dense single-segment mode */) {
    const int P = ctx->P, C = ctx->C;
    cudaStream_t s = ctx->stream;
    // which cluster counts go where
    uint32_t mask_pairs = 0, mask_prog = 0;
    bool irregular_any = false;
    for (int K = 1; K <= KMAX; ++K) {
        if (!ctx->k_present[K] || (host_seg && host_seg->K != K)) continue;
        if (precision == BAMSE_FP32 && ctx->geom_r[K] > 0) { mask_pairs |= 1u << K; irregular_any |= ctx->pairs[K].n_irregular > 0; }
        else mask_prog |= 1u << K;
    }
    if (precision == BAMSE_FP32 && (mask_prog || irregular_any)) { const int rc = ensure_interp(ctx); if (rc) return rc; }
    if (precision == BAMSE_FP32 && ctx->use_framed < 0 && !ctx->probing) {
        const char* f = getenv("BAMSE_FRAMED");
        if (f) ctx->use_framed = atoi(f) ? 1 : 0;
        else if (host_seg || mask_pairs || !ctx->owner.empty()) ctx->use_framed = 0;   // dense debug ranges / pair path / plan: no probe
        else { int rc = probe_fast_variant(ctx); if (rc) return rc; }
    }
    const int pi = precision == BAMSE_FP64 ? 1 : (precision == BAMSE_FP32 ? 0 : 2);
    const bool framed = precision == BAMSE_FP32 && ctx->use_framed == 1;
    if (precision == BAMSE_FP32) ctx->grid[pi] = enumerate_fast_grid_size(ctx->device, topk, framed, ctx->slots_needed);
    else if (ctx->grid[pi] == 0) ctx->grid[pi] = enumerate_grid_size(precision, ctx->device);
    const int grid = ctx->grid[pi];
    if (mask_pairs && ctx->grid_pairs == 0) ctx->grid_pairs = enumerate_pairs_grid_size(ctx->device);
    // work granularity of the program kernels: ~32 chunks per worker (warp for fp32, CTA for the check mode), 1..16 trees
    const int workers = enumerate_workers(precision, grid);
    const int workers2 = mask_pairs ? enumerate_pairs_workers(ctx->grid_pairs) : 0;
    uint64_t trees_prog = 0, trees_pairs = 0, irr_cap = 0;
    std::vector<unsigned long long> flat_off((size_t)P + 1, 0);
    if (host_seg) {
        if (mask_pairs) { trees_pairs = host_seg->n_trees; irr_cap = std::min<uint64_t>(host_seg->n_trees, ctx->pairs[host_seg->K].n_irregular); }
        else trees_prog = host_seg->n_trees;
        for (int p = 0; p <= P; ++p) flat_off[p] = p > host_seg->p ? irr_cap : 0;
    } else {
        for (int p = 0; p < P; ++p) {
            flat_off[p] = irr_cap;
            for (int c = 0; c < C; ++c) {
                const int K = ctx->K_of_c[c];
                const uint64_t T = ipow_u64(K, K - 1);
                const int own = ctx->owner.empty() ? -1 : ctx->owner[(size_t)p * C + c];
                // upper bound of this rank's trees of (p, c): an interleaved share, the whole range, or nothing
                const uint64_t mine = own < 0 ? (T + ctx->shard_world - 1) / ctx->shard_world : (own == ctx->shard_rank ? T : 0);
                if ((mask_pairs >> K) & 1u) {
                    // a split (p, c) is walked in full by the pair kernel of every rank (node-set sharding)
                    trees_pairs += (own < 0 && !ctx->owner.empty()) ? T : mine;
                    irr_cap += std::min<uint64_t>(mine + 1, ctx->pairs[K].n_irregular);
                } else trees_prog += mine;
            }
        }
        flat_off[P] = irr_cap;
        trees_prog += (uint64_t)P * C;
    }
    const int chunk = (int)std::min<uint64_t>(16, std::max<uint64_t>(1, trees_prog / ((uint64_t)workers * 32)));
    const int chunk2 = 32;                           // one tree slot per lane
    // sample-major pair path (the warps in flight share one sample's tables, L2 resident): when every pair cluster
    // count is small enough for that to matter (K <= 8), nothing exits early and the per-sample results fit
    bool sample_major = mask_pairs != 0 && ctx->M > 1 && !(ctx->early_exit && !host_seg && topk > 0) && (mask_pairs >> 9) == 0;
    const uint64_t part_slots = (trees_pairs / chunk2 + (uint64_t)P * C + 1) * chunk2;
    if (sample_major && part_slots * (uint64_t)ctx->M * sizeof(float) > (8ull << 30)) sample_major = false;
    if (const char* e = getenv("BAMSE_SAMPLE_MAJOR")) sample_major = sample_major && atoi(e) != 0;
    const int reps2 = sample_major ? ctx->M : 1;
    const uint64_t chunks_prog = trees_prog / chunk + (uint64_t)P * C + 1;
    const uint64_t chunks_pairs = trees_pairs / chunk2 + (uint64_t)P * C + 1;
    uint64_t cand_lists = std::min<uint64_t>((uint64_t)workers * P, chunks_prog + workers);
    if (mask_pairs) cand_lists += std::min<uint64_t>((uint64_t)workers2 * P, chunks_pairs + workers2) +
                                  std::min<uint64_t>((uint64_t)workers * P, irr_cap + workers);
    const uint64_t cand_cap = cand_lists * (uint64_t)(topk > 0 ? topk : 0);
    BAMSE_CUDA_TRY(ctx, ctx->d_segs.ensure(sizeof(Segment) * P * C));
    BAMSE_CUDA_TRY(ctx, ctx->d_segs2.ensure(sizeof(Segment) * P * C));
    const bool fresh_scalars = ctx->d_scalars.p == nullptr;
    BAMSE_CUDA_TRY(ctx, ctx->d_scalars.ensure(sizeof(uint64_t) * 16));
    uint64_t* scal = ctx->d_scalars.as<uint64_t>();
    // [0] chunks (programs) [1] trees [2] next chunk [3] bag cursor [4..8) statistics (accumulate until read)
    // [8] samples evaluated [10] chunks (pairs) [11] trees [12] next chunk (pairs) [13] next chunk (flat) [14] flat overflow
    if (fresh_scalars) BAMSE_CUDA_TRY(ctx, cudaMemsetAsync(scal, 0, sizeof(uint64_t) * 16, s));
    else {
        BAMSE_CUDA_TRY(ctx, cudaMemsetAsync(scal, 0, sizeof(uint64_t) * 4, s));
        BAMSE_CUDA_TRY(ctx, cudaMemsetAsync(scal + 9, 0, sizeof(uint64_t) * 7, s));   // [9] trees taken by the pair kernel
    }
    if (topk > 0) BAMSE_CUDA_TRY(ctx, ctx->d_cand.ensure(sizeof(bamse_record) * (size_t)cand_cap));
    unsigned long long* flat_meta = nullptr;         // [P + 1] region offsets, [P] counts, [P + 1] running sums
    if (mask_pairs) {
        BAMSE_CUDA_TRY(ctx, ctx->d_flat.ensure(sizeof(unsigned long long) * std::max<uint64_t>(irr_cap, 1)));
        BAMSE_CUDA_TRY(ctx, ctx->d_flat_meta.ensure(sizeof(unsigned long long) * (3 * (size_t)P + 2)));
        flat_meta = ctx->d_flat_meta.as<unsigned long long>();
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(flat_meta, flat_off.data(), sizeof(unsigned long long) * (P + 1), cudaMemcpyHostToDevice, s));
        BAMSE_CUDA_TRY(ctx, cudaMemsetAsync(flat_meta + P + 1, 0, sizeof(unsigned long long) * (2 * (size_t)P + 1), s));
        BAMSE_CUDA_TRY(ctx, ctx->d_lists2.ensure(sizeof(bamse_record) * (size_t)workers2 * (size_t)(topk > 0 ? topk : 1)));
        if (sample_major) BAMSE_CUDA_TRY(ctx, ctx->d_part.ensure(sizeof(float) * part_slots * (size_t)ctx->M));
    }
    const int32_t* d_owner = ctx->owner.empty() ? nullptr : ctx->d_owner.as<int32_t>();
    EnumWork w;
    memset(&w, 0, sizeof w);
    if (host_seg) {
        const Segment seg = *host_seg;
        const uint64_t chunks = (seg.n_trees + (mask_pairs ? chunk2 : chunk) - 1) / (mask_pairs ? chunk2 : chunk) * (mask_pairs ? reps2 : 1);
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(mask_pairs ? ctx->d_segs2.p : ctx->d_segs.p, &seg, sizeof seg, cudaMemcpyHostToDevice, s));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(mask_pairs ? scal + 10 : scal, &chunks, sizeof chunks, cudaMemcpyHostToDevice, s));
        BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));
        w.n_segs = 1;
        ctx->last_evals = (int64_t)seg.n_trees;
    } else {
        if (mask_prog || !mask_pairs) {
            launch_build_segments(P, C, ctx->d_K.as<int32_t>(), d_const, d_thr, d_begin, d_end, slack_abs, slack_rel,
                                  chunk, 1, ctx->shard_rank, ctx->shard_world, mask_pairs ? mask_prog : 0xffffffffu, 0, d_owner,
                                  ctx->d_segs.as<Segment>(), scal, scal + 1, s);
            ctx->launches++;
        }
        if (mask_pairs) {
            launch_build_segments(P, C, ctx->d_K.as<int32_t>(), d_const, d_thr, d_begin, d_end, slack_abs, slack_rel,
                                  chunk2, reps2, ctx->shard_rank, ctx->shard_world, mask_pairs, d_owner ? 1 : 0, d_owner,
                                  ctx->d_segs2.as<Segment>(), scal + 10, scal + 11, s);
            ctx->launches++;
        }
        w.n_segs = P * C;
    }
    w.topk = topk;
    w.cand = ctx->d_cand.as<bamse_record>();
    w.cand_count = reinterpret_cast<unsigned long long*>(scal + 3);
    w.cand_cap = cand_cap;
    ctx->last_cand_cap = cand_cap;
    w.dense_score = d_dense_score;
    w.dense_logscre = d_dense_logscre;
    w.stats = reinterpret_cast<unsigned long long*>(scal + 4);
    w.samples_done = reinterpret_cast<unsigned long long*>(scal + 8);
    w.trees_done = reinterpret_cast<unsigned long long*>(scal + 9);
    w.lists = nullptr;
    if (precision == BAMSE_FP32) {
        BAMSE_CUDA_TRY(ctx, ctx->d_lists.ensure(sizeof(bamse_record) * (size_t)workers * (size_t)(topk > 0 ? topk : 1)));
        w.lists = ctx->d_lists.as<bamse_record>();
    }
    w.lists2 = ctx->d_lists2.as<bamse_record>();
    w.prog_tabs = (precision == BAMSE_FP32 && ctx->have_progtabs) ? ctx->d_progtabs.as<const FastProgram*>() : nullptr;
    // early exit only where pruned trees cannot matter: top-k enumerations (not dense ranges, not probes)
    w.early_exit = (ctx->early_exit && !host_seg && !ctx->probing && topk > 0) ? 1 : 0;
    if (mask_pairs) {
        w.flat = ctx->d_flat.as<unsigned long long>();
        w.flat_off = flat_meta;
        w.flat_cnt = flat_meta + P + 1;
        w.flat_prefix = flat_meta + 2 * P + 1;
        w.flat_overflow = reinterpret_cast<unsigned long long*>(scal + 14);
    }
    // tail splitting: when a worker gets fewer than ~48 single-tree chunks, issue the last `workers`
    // trees in 2 (or, below 16 chunks per worker, 4) parts over the samples; always for the flat list
    const char* no_split = getenv("BAMSE_SPLIT");      // BAMSE_SPLIT=0 disables (A/B measurements)
    const bool split_ok = precision == BAMSE_FP32 && !ctx->probing && ctx->M >= 4 && !(no_split && atoi(no_split) == 0);
    const bool split_prog = split_ok && chunk == 1 && trees_prog < (uint64_t)workers * 48;
    const bool split_flat = split_ok && irregular_any;
    if (split_prog || split_flat) {
        const size_t bytes = (sizeof(double) + sizeof(int)) * (size_t)workers * 2;
        BAMSE_CUDA_TRY(ctx, ctx->d_split.ensure(bytes));
        BAMSE_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_split.p, 0, bytes, s));
    }
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!ctx->ev_free.empty()) { ev = ctx->ev_free.back(); ctx->ev_free.pop_back(); }
    else {
        BAMSE_CUDA_TRY(ctx, cudaEventCreate(&ev.first));
        BAMSE_CUDA_TRY(ctx, cudaEventCreate(&ev.second));
    }
    BAMSE_CUDA_TRY(ctx, cudaEventRecord(ev.first, s));
    if (mask_pairs) {
        EnumWork wp = w;
        wp.segs = ctx->d_segs2.as<Segment>();
        wp.chunk = chunk2;
        wp.total_chunks = scal + 10;
        wp.next_chunk = reinterpret_cast<unsigned long long*>(scal + 12);
        wp.part = ctx->d_part.as<float>();
        // chunks per visit of the work counter: ~64 visits per warp
        wp.grab = (int)std::min<uint64_t>(64, std::max<uint64_t>(1, chunks_pairs * (uint64_t)reps2 / ((uint64_t)workers2 * 64)));
        if (sample_major) {
            launch_enumerate_pairs(ctx->view(), ctx->fast(), wp, ctx->grid_pairs, 1, s);
            EnumWork wr = wp;                          // second pass: its own chunk counter
            wr.next_chunk = reinterpret_cast<unsigned long long*>(scal + 15);
            launch_enumerate_pairs(ctx->view(), ctx->fast(), wr, ctx->grid_pairs, 2, s);
            ctx->launches += 2;
        } else {
            launch_enumerate_pairs(ctx->view(), ctx->fast(), wp, ctx->grid_pairs, 0, s);
            ctx->launches++;
        }
        if (irregular_any) {
            launch_flat_prefix(P, wp.flat_cnt, wp.flat_off, wp.flat_prefix, s);
            EnumWork wf = wp;
            wf.flat_mode = 1;
            wf.chunk = 1;
            wf.next_chunk = reinterpret_cast<unsigned long long*>(scal + 13);
            wf.split_ways = 1;
            if (split_flat) {
                wf.split_ways = ctx->M >= 8 ? 4 : 2;
                wf.n_split = (unsigned long long)workers;
                wf.split_acc = ctx->d_split.as<double>() + workers;
                wf.split_cnt = reinterpret_cast<int*>(ctx->d_split.as<double>() + 2 * (size_t)workers) + workers;
            }
            launch_enumerate_fast(ctx->view(), ctx->fast(), wf, grid, framed, ctx->slots_needed, s);
            ctx->launches += 2;
        }
    }
    if (mask_prog || !mask_pairs) {
        w.segs = ctx->d_segs.as<Segment>();
        w.chunk = chunk;
        w.total_chunks = scal;
        w.next_chunk = reinterpret_cast<unsigned long long*>(scal + 2);
        w.flat = nullptr;
        w.split_ways = 1;
        if (split_prog) {
            w.split_ways = (ctx->M >= 8 && trees_prog < (uint64_t)workers * 16) ? 4 : 2;
            w.n_split = (unsigned long long)workers;
            w.split_acc = ctx->d_split.as<double>();
            w.split_cnt = reinterpret_cast<int*>(ctx->d_split.as<double>() + 2 * (size_t)workers);
        }
        if (precision == BAMSE_FP32) launch_enumerate_fast(ctx->view(), ctx->fast(), w, grid, framed, ctx->slots_needed, s);
        else launch_enumerate(precision, ctx->view(), w, grid, s);
        ctx->launches++;
    }
    BAMSE_CUDA_TRY(ctx, cudaEventRecord(ev.second, s));
    ctx->ev_used.push_back(ev);
    if (ctx->ev_used.size() > 256) {
        ctx->ev_free.push_back(ctx->ev_used.front());
        ctx->ev_used.erase(ctx->ev_used.begin());
    }
    BAMSE_CUDA_TRY(ctx, cudaGetLastError());
    if (topk > 0 && d_out) {
        launch_merge_topk(ctx->d_cand.as<bamse_record>(), reinterpret_cast<unsigned long long*>(scal + 3), cand_cap, P,
                          topk, d_out, s);
        ctx->launches++;
        BAMSE_CUDA_TRY(ctx, cudaGetLastError());
    }
    return BAMSE_OK;
}

int bamse_score_range(bamse_ctx* ctx, int p, int c, uint64_t tree_begin, int64_t n, int precision, double* out_score,
                      double* out_logscre) try {
    CTX_CHECK(ctx);
    if (!ctx->have_problem) return ctx->fail(BAMSE_ESTATE, "bamse_score_range: no problem uploaded");
    if (p < 0 || p >= ctx->P || c < 0 || c >= ctx->C || n < 0)
        return ctx->fail(BAMSE_EINVAL, "bamse_score_range: p/c/n out of range");
    if (!valid_precision(precision))
        return ctx->fail(BAMSE_EINVAL, "bamse_score_range: bad precision");
    const int K = ctx->K_of_c[c];
    const uint64_t T = ipow_u64(K, K - 1);
    if (tree_begin > T || (uint64_t)n > T - tree_begin)
        return ctx->fail(BAMSE_EINVAL, "bamse_score_range: range exceeds K^(K-1)=%llu", (unsigned long long)T);
    if (n == 0) return BAMSE_OK;
    if (precision == BAMSE_FP32 && !ctx->owner.empty() && ctx->owner[(size_t)p * ctx->C + c] != ctx->shard_rank)
        return ctx->fail(BAMSE_ESTATE, "bamse_score_range: under the sharding plan this rank does not hold all tables of "
                                       "(p=%d, c=%d); upload the problem unsharded for dense ranges", p, c);
    BAMSE_CUDA_TRY(ctx, ctx->d_dense0.ensure(sizeof(double) * n));
    BAMSE_CUDA_TRY(ctx, ctx->d_dense1.ensure(sizeof(double) * n));
    Segment seg;
    memset(&seg, 0, sizeof seg);
    seg.p = p; seg.c = c; seg.K = K; seg.stride = 1;
    seg.tree_begin = tree_begin; seg.n_trees = (uint64_t)n;
    seg.first_chunk = 0; seg.clust_const = 0.0; seg.threshold = -INFINITY;
    seg.whole = 0; seg.pad = 0;
    int rc = enumerate_dev(ctx, nullptr, nullptr, nullptr, nullptr, 0, precision, 0.0, 0.0, nullptr,
                           ctx->d_dense0.as<double>(), ctx->d_dense1.as<double>(), &seg);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    if (out_score) BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(out_score, ctx->d_dense0.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    if (out_logscre) BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(out_logscre, ctx->d_dense1.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return BAMSE_OK;
} catch (...) { return on_exception(ctx, "bamse_score_range"); }

int bamse_enumerate_topk_dev(bamse_ctx* ctx, const double* d_clust_const, const double* d_threshold,
                             const uint64_t* d_tree_begin, const uint64_t* d_tree_end, int topk, int precision,
                             double slack, bamse_record* d_out_records) try {
    CTX_CHECK(ctx);
    if (!ctx->have_problem) return ctx->fail(BAMSE_ESTATE, "bamse_enumerate_topk_dev: no problem uploaded");
    if (topk < 1 || topk > BAMSE_TOPK_MAX || !d_out_records)
        return ctx->fail(BAMSE_EINVAL, "bamse_enumerate_topk_dev: topk must be in [1,%d], d_out_records non-NULL", BAMSE_TOPK_MAX);
    if (!valid_precision(precision))
        return ctx->fail(BAMSE_EINVAL, "bamse_enumerate_topk_dev: bad precision");
    // eval count (host arithmetic; ranges live on the device, so count full ranges unless given)
    ctx->last_evals = -1;
    return enumerate_dev(ctx, d_clust_const, d_threshold, d_tree_begin, d_tree_end, topk, precision, slack, 0.0,
                         d_out_records, nullptr, nullptr, nullptr);
} catch (...) { return on_exception(ctx, "bamse_enumerate_topk_dev"); }

int bamse_merge_topk_dev(bamse_ctx* ctx, const bamse_record* d_lists, int n_lists, int P, int topk,
                         bamse_record* d_out_records) try {
    CTX_CHECK(ctx);
    if (!d_lists || !d_out_records || n_lists < 1 || P < 1 || topk < 1 || topk > BAMSE_TOPK_MAX)
        return ctx->fail(BAMSE_EINVAL, "bamse_merge_topk_dev: bad arguments");
    launch_merge_topk(d_lists, nullptr, (unsigned long long)n_lists * P * topk, P, topk, d_out_records, ctx->stream);
    ctx->launches++;
    BAMSE_CUDA_TRY(ctx, cudaGetLastError());
    return BAMSE_OK;
} catch (...) { return on_exception(ctx, "bamse_merge_topk_dev"); }

int bamse_score_topk(bamse_ctx* ctx, const double* clust_const, const double* threshold, const uint64_t* tree_begin,
                     const uint64_t* tree_end, int topk, int precision, bamse_record* out_records, int8_t* out_parent,
                     int32_t* out_count) try {
    CTX_CHECK(ctx);
    if (!ctx->have_problem) return ctx->fail(BAMSE_ESTATE, "bamse_score_topk: no problem uploaded");
    if (!clust_const || !out_records || topk < 1 || topk > BAMSE_TOPK_MAX)
        return ctx->fail(BAMSE_EINVAL, "bamse_score_topk: clust_const/out_records NULL or topk not in [1,%d]", BAMSE_TOPK_MAX);
    if (!valid_precision(precision))
        return ctx->fail(BAMSE_EINVAL, "bamse_score_topk: bad precision");
    if ((tree_begin == nullptr) != (tree_end == nullptr))
        return ctx->fail(BAMSE_EINVAL, "bamse_score_topk: tree_begin and tree_end must both be given or both NULL");
    const int P = ctx->P, C = ctx->C, Kmax = ctx->Kmax;
    cudaStream_t s = ctx->stream;
    BAMSE_CUDA_TRY(ctx, ctx->d_const.ensure(sizeof(double) * P * C));
    BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_const.p, clust_const, sizeof(double) * P * C, cudaMemcpyHostToDevice, s));
    if (threshold) {
        BAMSE_CUDA_TRY(ctx, ctx->d_thr.ensure(sizeof(double) * P * C));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_thr.p, threshold, sizeof(double) * P * C, cudaMemcpyHostToDevice, s));
    }
    int64_t evals = 0, evals_prog = 0;
    bool evals_split_pending = false;
    for (int p = 0; p < P; ++p)
        for (int c = 0; c < C; ++c) {
            const int K = ctx->K_of_c[c];
            const uint64_t T = ipow_u64(K, K - 1);
            uint64_t b = tree_begin ? tree_begin[c] : 0, e = tree_end ? std::min<uint64_t>(tree_end[c], T) : T;
            const int own = ctx->owner.empty() ? -1 : ctx->owner[(size_t)p * C + c];
            int64_t n = 0;
            if (own >= 0) { if (own == ctx->shard_rank && b < e) n = (int64_t)(e - b); }
            else if (b < e && e - b > (uint64_t)ctx->shard_rank)
                n = (int64_t)((e - b - ctx->shard_rank + ctx->shard_world - 1) / ctx->shard_world);
            evals += n;
            // trees that go through the pair kernel are counted there when some (p, c) is split by node sets
            if (!(precision == BAMSE_FP32 && ctx->geom_r[K] > 0)) evals_prog += n;
            if (precision == BAMSE_FP32 && ctx->geom_r[K] > 0 && own < 0 && !ctx->owner.empty()) evals_split_pending = true;
        }
    if (tree_begin) {
        BAMSE_CUDA_TRY(ctx, ctx->d_begin.ensure(sizeof(uint64_t) * C));
        BAMSE_CUDA_TRY(ctx, ctx->d_end.ensure(sizeof(uint64_t) * C));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_begin.p, tree_begin, sizeof(uint64_t) * C, cudaMemcpyHostToDevice, s));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_end.p, tree_end, sizeof(uint64_t) * C, cudaMemcpyHostToDevice, s));
    }
    // Preselection + verified cut.  The enumeration kernel ranks by ITS OWN arithmetic (fp32, or the check
    // mode's device-side sums); the returned ranking must be the one of the fp64 check-mode scores of the
    // list-scoring kernel.  So: select the best `ksel` > topk per parameter set (thresholds relaxed by a slack
    // that covers the selection arithmetic's error), re-score exactly that short list in fp64, apply the exact
    // threshold, rank -- and then PROVE that nothing outside the list could belong to the top `topk`: every
    // unselected tree has a selection total <= the worst selected one, T_last, and an fp64 total of at most
    // T_last + B with B = rel * max_c |T_last - const[p][c]| + abs (the error bound of the selection
    // arithmetic on a score of that magnitude; the bound of a tree further down grows slower than its
    // distance).  The cut is safe when the list was not full (everything that qualified is in it) or when
    // the topk-th re-scored total exceeds T_last + B; otherwise ksel is doubled and the enumeration repeated
    // (near-degenerate clusterings: many trees tied within rounding).  At the internal cap the result is
    // returned best-effort and bamse_last_topk_verified() reports 0.
    const double sel_rel = precision == BAMSE_FP64 ? 1e-11 : 2e-6;      // error bound of the selection arithmetic:
    const double sel_abs = precision == BAMSE_FP64 ? 1e-9 : 1e-4;       // relative to |score|, plus absolute
    const double slack_abs = precision == BAMSE_FP64 ? 1e-9 : 1e-3;     // threshold relaxation (>= the bound)
    const double slack_rel = precision == BAMSE_FP64 ? 1e-11 : 1e-5;
    int ksel = std::min(TOPK_INTERNAL, topk + 8);
    std::vector<bamse_record> recs;
    ctx->last_verified = 1;
    for (;;) {
        BAMSE_CUDA_TRY(ctx, ctx->d_records.ensure(sizeof(bamse_record) * P * ksel));
        int rc = enumerate_dev(ctx, ctx->d_const.as<double>(), threshold ? ctx->d_thr.as<double>() : nullptr,
                               tree_begin ? ctx->d_begin.as<uint64_t>() : nullptr,
                               tree_begin ? ctx->d_end.as<uint64_t>() : nullptr, ksel, precision, slack_abs, slack_rel,
                               ctx->d_records.as<bamse_record>(), nullptr, nullptr, nullptr);
        if (rc) return rc;
        ctx->last_evals = evals;
        recs.assign((size_t)P * ksel, bamse_record());
        uint64_t bag[2] = {0, 0};                     // scalars [3] = records appended to the bag, [14] = flat-list overflow
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(recs.data(), ctx->d_records.p, sizeof(bamse_record) * P * ksel, cudaMemcpyDeviceToHost, s));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(bag, ctx->d_scalars.as<uint64_t>() + 3, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        BAMSE_CUDA_TRY(ctx, cudaMemcpyAsync(bag + 1, ctx->d_scalars.as<uint64_t>() + 14, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        BAMSE_CUDA_TRY(ctx, cudaStreamSynchronize(s));
        if (bag[0] > ctx->last_cand_cap)
            return ctx->fail(BAMSE_EINTERNAL, "bamse_score_topk: candidate bag overflow (%llu records, capacity %llu)",
                             (unsigned long long)bag[0], (unsigned long long)ctx->last_cand_cap);
        if (bag[1] != 0)
            return ctx->fail(BAMSE_EINTERNAL, "bamse_score_topk: the list of irregular trees overflowed its capacity");
        if (precision == BAMSE_FP32 && evals_split_pending) {
            // the share of a split (p, c) is decided by the pair kernel (node sets): take its own count
            uint64_t taken = 0;
            BAMSE_CUDA_TRY(ctx, cudaMemcpy(&taken, ctx->d_scalars.as<uint64_t>() + 9, sizeof taken, cudaMemcpyDeviceToHost));
            ctx->last_evals = evals_prog + (int64_t)taken;
        }
        // worst selected total and list fill per parameter set, before the re-score reorders anything
        std::vector<double> t_last(P, -INFINITY);
        std::vector<int> n_sel(P, 0);
        for (int p = 0; p < P; ++p)
            for (int j = 0; j < ksel; ++j) {
                const bamse_record& r = recs[(size_t)p * ksel + j];
                if (r.clust >= 0) { t_last[p] = r.total; n_sel[p]++; }
            }
        // fp64 check-mode re-score of the short list, exact threshold, final ranking
        std::vector<ScoreItem> items;
        std::vector<size_t> where;
        std::vector<double> logscre;
        for (size_t i = 0; i < recs.size(); ++i) {
            if (recs[i].clust < 0) continue;
            ScoreItem it;
            memset(&it, 0, sizeof it);
            it.p = recs[i].param; it.c = recs[i].clust;
            const int K = ctx->K_of_c[it.c];
            it.k = (int8_t)K;
            decode_tree(K, recs[i].tree, it.parent);
            for (int v = 0; v < K; ++v) it.nodes[v] = (int8_t)v;
            TreeProgram prog;
            build_program(K, it.parent, it.nodes, &prog);
            logscre.push_back(log((double)prog.scre));
            items.push_back(it);
            where.push_back(i);
        }
        std::vector<double> sc(items.size());
        rc = score_items_host(ctx, items, BAMSE_FP64, sc.data());
        if (rc) return rc;
        for (size_t j = 0; j < items.size(); ++j) {
            bamse_record& r = recs[where[j]];
            r.score = sc[j] + logscre[j];
            r.total = r.score + clust_const[(size_t)r.param * C + r.clust];
            if (threshold) {
                const double th = threshold[(size_t)r.param * C + r.clust];
                if (!(r.score >= th - 1e-12 * fabs(th))) r.clust = -1;
            }
        }
        bool verified = true;
        for (int p = 0; p < P; ++p) {
            bamse_record* b = recs.data() + (size_t)p * ksel;
            std::stable_sort(b, b + ksel, [](const bamse_record& x, const bamse_record& y) {
                if ((x.clust < 0) != (y.clust < 0)) return y.clust < 0;
                if (x.clust < 0) return false;
                if (x.total != y.total) return x.total > y.total;
                if (x.clust != y.clust) return x.clust < y.clust;
                return x.tree < y.tree;
            });
            if (n_sel[p] < ksel) continue;            // the list holds every tree that passed the relaxed threshold
            double spread = 0.0;
            for (int c = 0; c < C; ++c) spread = std::max(spread, fabs(t_last[p] - clust_const[(size_t)p * C + c]));
            const double bound = sel_rel * spread + sel_abs;
            const bamse_record& kth = b[topk - 1];    // ksel > topk here, or ksel == TOPK_INTERNAL
            if (!(kth.clust >= 0 && kth.total > t_last[p] + bound)) verified = false;
        }
        ctx->last_ksel = ksel;
        if (verified) break;
        if (ksel >= TOPK_INTERNAL) { ctx->last_verified = 0; break; }
        ksel = std::min(TOPK_INTERNAL, 2 * ksel);
    }
    for (int p = 0; p < P; ++p) {
        int cnt = 0;
        for (int j = 0; j < topk; ++j) {
            bamse_record r = recs[(size_t)p * ksel + j];
            if (r.clust < 0) { r.total = -INFINITY; r.score = -INFINITY; r.clust = -1; r.param = p; r.tree = 0; }
            else ++cnt;
            out_records[(size_t)p * topk + j] = r;
            if (out_parent) {
                int8_t* par = out_parent + ((size_t)p * topk + j) * Kmax;
                for (int v = 0; v < Kmax; ++v) par[v] = -2;
                if (r.clust >= 0) decode_tree(ctx->K_of_c[r.clust], r.tree, par);
            }
        }
        if (out_count) out_count[p] = cnt;
    }
    return BAMSE_OK;
} catch (...) { return on_exception(ctx, "bamse_score_topk"); }

}  // extern "C"
