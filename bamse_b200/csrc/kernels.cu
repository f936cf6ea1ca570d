// Kernels of libbamse_cuda for sm_100a: table construction, work segmentation,
// log-domain enumeration / list scoring (fp64 check mode + fp32), top-k merge.
#include "kernels.cuh"
#include "score_logdomain.cuh"
#include "tree_decode.cuh"

namespace bamse {

// ------------------------------------------------------------------------- tables
// D[p][c][m][k][x] = logC + log(cf_x)*a + log(1-cf_x)*b     (clustering.py:23), fp64 with the
// reference's operation order (no FMA contraction), plus an fp32 rounding of it.
__global__ void __launch_bounds__(256) k_tables(ProblemView pv, const double* __restrict__ var,
                                                const double* __restrict__ ref, const double* __restrict__ logc,
                                                const double* __restrict__ logcf, const double* __restrict__ log1mcf,
                                                double* __restrict__ D64, float* __restrict__ D32) {
    const int row = blockIdx.x;                 // ((p*C + c)*M + m)*Kmax + k
    const int k = row % pv.Kmax;
    const int m = (row / pv.Kmax) % pv.M;
    const int c = (row / (pv.Kmax * pv.M)) % pv.C;
    const int p = row / (pv.Kmax * pv.M * pv.C);
    const size_t cnt = ((size_t)c * pv.Kmax + k) * pv.M + m;
    const bool live = k < pv.K_of_c[c];
    const double a = live ? var[cnt] : 0.0, b = live ? ref[cnt] : 0.0, lc = live ? logc[cnt] : 0.0;
    for (int x = threadIdx.x; x < L; x += blockDim.x) {
        const double t1 = __dmul_rn(logcf[p * L + x], a);
        const double t2 = __dmul_rn(log1mcf[p * L + x], b);
        const double y = __dadd_rn(__dadd_rn(lc, t1), t2);
        D64[(size_t)row * L + x] = y;
        D32[(size_t)row * L + x] = (float)y;
    }
}

void launch_tables(const ProblemView& pv, const double* d_var, const double* d_ref, const double* d_logc,
                   const double* d_logcf, const double* d_log1mcf, double* D64, float* D32, cudaStream_t s) {
    const int rows = pv.P * pv.C * pv.M * pv.Kmax;
    k_tables<<<rows, 256, 0, s>>>(pv, d_var, d_ref, d_logc, d_logcf, d_log1mcf, D64, D32);
}

// ----------------------------------------------------------------------- segments
__global__ void __launch_bounds__(256) k_build_segments(int P, int C, const int32_t* __restrict__ K_of_c,
                                                        const double* __restrict__ cconst, const double* __restrict__ thr,
                                                        const uint64_t* __restrict__ begin, const uint64_t* __restrict__ end,
                                                        double slack_abs, double slack_rel, int chunk, int reps, int shard_rank,
                                                        int shard_world, uint32_t k_mask, int split_full, const int32_t* __restrict__ owner,
                                                        Segment* __restrict__ segs,
                                                        uint64_t* __restrict__ total_chunks, uint64_t* __restrict__ total_trees) {
    // one thread per (p, c) fills its segment, then thread 0 turns chunk counts into prefix sums
    const int n = P * C;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int p = i / C, c = i - p * C;
        const int K = K_of_c[c];
        const uint64_t T = ipow_u64(K, K - 1);
        uint64_t b = begin ? begin[c] : 0, e = end ? end[c] : T;
        if (e > T) e = T;
        if (b > e) b = e;
        Segment s;
        s.p = p; s.c = c; s.K = K;
        // sharding plan: owner < 0 (or no plan) = interleaved over all ranks (indices b + rank, b + rank + world,
        // ... below e), owner == rank = this rank scores the whole range, any other owner = nothing here
        const int own = owner ? owner[i] : -1;
        const int rank = (own < 0 && !split_full) ? shard_rank : 0, world = (own < 0 && !split_full) ? shard_world : 1;
        const uint64_t span = (own >= 0 && own != shard_rank) || !((k_mask >> K) & 1u) ? 0 : e - b;
        s.stride = world;
        s.tree_begin = b + (uint64_t)rank;
        s.n_trees = span > (uint64_t)rank ? (span - rank + world - 1) / world : 0;
        s.first_chunk = (s.n_trees + (uint64_t)chunk - 1) / (uint64_t)chunk * (uint64_t)reps;   // count for now
        s.whole = (s.stride == 1 && s.tree_begin == 0 && s.n_trees == T) ? 1 : 0;
        s.pad = 0;
        s.clust_const = cconst ? cconst[i] : 0.0;
        double th = thr ? thr[i] : -CUDART_INF;
        if (th > -CUDART_INF) th -= slack_abs + slack_rel * fabs(th);
        s.threshold = th;
        segs[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t chunks = 0, trees = 0;
        for (int i = 0; i < n; ++i) {
            const uint64_t cnt = segs[i].first_chunk;
            segs[i].first_chunk = chunks;
            chunks += cnt;
            trees += segs[i].n_trees;
        }
        *total_chunks = chunks;
        *total_trees = trees;
    }
}

void launch_build_segments(int P, int C, const int32_t* d_K_of_c, const double* d_const, const double* d_thr,
                           const uint64_t* d_begin, const uint64_t* d_end, double slack_abs, double slack_rel,
                           int chunk, int reps, int shard_rank, int shard_world, uint32_t k_mask, int split_full,
                           const int32_t* d_owner,
                           Segment* d_segs, uint64_t* d_total_chunks, uint64_t* d_total_trees, cudaStream_t s) {
    k_build_segments<<<1, 256, 0, s>>>(P, C, d_K_of_c, d_const, d_thr, d_begin, d_end, slack_abs, slack_rel, chunk,
                                      reps, shard_rank, shard_world, k_mask, split_full, d_owner, d_segs, d_total_chunks, d_total_trees);
}

// --------------------------------------------------------------------- top-k helper
// strict total order: total desc, then clust asc, then tree asc (param is per list)
__device__ __forceinline__ bool rec_better(const bamse_record& a, const bamse_record& b) {
    if (a.total != b.total) return a.total > b.total;
    if (a.clust != b.clust) return a.clust < b.clust;
    return a.tree < b.tree;
}

__device__ inline void topk_insert(bamse_record* list, int* n, int k, const bamse_record& r) {
    int cnt = *n;
    if (cnt == k && !rec_better(r, list[k - 1])) return;
    int i = cnt < k ? cnt : k - 1;
    while (i > 0 && rec_better(r, list[i - 1])) { list[i] = list[i - 1]; --i; }
    list[i] = r;
    if (cnt < k) *n = cnt + 1;
}

// ---------------------------------------------------------------------- enumerate
template <typename T>
__global__ void __launch_bounds__(LD_THREADS) k_enumerate_logdomain(ProblemView pv, EnumWork w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* pool = reinterpret_cast<T*>(smem_raw);
    __shared__ T scratch[LD_THREADS / 32];
    __shared__ TreeProgram prog;
    __shared__ unsigned long long s_chunk;
    __shared__ bamse_record s_top[BAMSE_TOPK_INTERNAL];
    __shared__ int s_ntop;
    const int t = threadIdx.x;
    if (t == 0) s_ntop = 0;
    int cur_p = -1;
    const uint64_t total = *w.total_chunks;
    for (;;) {
        __syncthreads();
        if (t == 0) s_chunk = atomicAdd(w.next_chunk, 1ull);
        __syncthreads();
        const uint64_t chunk = s_chunk;
        if (chunk >= total) break;
        int lo = 0, hi = w.n_segs - 1;              // last segment with first_chunk <= chunk and trees > 0
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (w.segs[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
        }
        const Segment seg = w.segs[lo];             // (empty segments share first_chunk with the next one;
                                                    //  the search lands on the last, i.e. the non-empty one)
        if (seg.p != cur_p) {
            if (t == 0 && cur_p >= 0 && s_ntop > 0) {
                const unsigned long long pos = atomicAdd(w.cand_count, (unsigned long long)s_ntop);
                if (pos + s_ntop <= w.cand_cap)
                    for (int i = 0; i < s_ntop; ++i) w.cand[pos + i] = s_top[i];
                s_ntop = 0;
            }
            cur_p = seg.p;
        }
        const uint64_t off = (chunk - seg.first_chunk) * (uint64_t)w.chunk;
        const uint64_t remain = seg.n_trees - off;
        const int n = remain < (uint64_t)w.chunk ? (int)remain : w.chunk;
        for (int i = 0; i < n; ++i) {
            const uint64_t tree = seg.tree_begin + (off + i) * (uint64_t)seg.stride;
            __syncthreads();
            if (t == 0) {
                int8_t parent[KMAX];
                decode_tree(seg.K, tree, parent);
                build_program(seg.K, parent, nullptr, &prog);
            }
            __syncthreads();
            double cut = -CUDART_INF;
            if (w.early_exit) {
                // what this tree needs: its clustering's threshold, or the CTA's k-th best so far
                double c0 = seg.threshold;
                if (w.topk > 0 && s_ntop == w.topk) c0 = fmax(c0, s_top[w.topk - 1].total - seg.clust_const);
                cut = c0 - log((double)prog.scre) - 1e-6 - 1e-9 * fabs(c0);
            }
            unsigned done = 0;
            const double score = eval_tree_logdomain<T>(pv, seg.p, seg.c, &prog, pool, scratch, 0, -1, cut, &done);
            if (t == 0) {
                if (w.samples_done) atomicAdd(w.samples_done, (unsigned long long)done);
                const double logscre = log((double)prog.scre);
                const double full = score + logscre;                 // clustering.py:182
                if (w.dense_score) w.dense_score[off + i] = score;
                if (w.dense_logscre) w.dense_logscre[off + i] = logscre;
                if (w.topk > 0 && full >= seg.threshold && full > -CUDART_INF) {   // clustering.py:146-147
                    bamse_record r;
                    r.total = full + seg.clust_const;                // clustering.py:183
                    r.score = full; r.clust = seg.c; r.param = seg.p; r.tree = tree;
                    topk_insert(s_top, &s_ntop, w.topk, r);
                }
            }
        }
    }
    if (t == 0 && cur_p >= 0 && s_ntop > 0) {
        const unsigned long long pos = atomicAdd(w.cand_count, (unsigned long long)s_ntop);
        if (pos + s_ntop <= w.cand_cap)
            for (int i = 0; i < s_ntop; ++i) w.cand[pos + i] = s_top[i];
    }
}

template <typename T> static size_t pool_bytes() { return (size_t)MAX_STACK * L * sizeof(T); }

int enumerate_grid_size(int precision, int device) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (precision == BAMSE_FP64)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_enumerate_logdomain<double>, LD_THREADS,
                                                      pool_bytes<double>());
    else
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_enumerate_logdomain<float>, LD_THREADS,
                                                      pool_bytes<float>());
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

void launch_enumerate(int precision, const ProblemView& pv, const EnumWork& w, int grid, cudaStream_t s) {
    if (precision == BAMSE_FP64)
        k_enumerate_logdomain<double><<<grid, LD_THREADS, pool_bytes<double>(), s>>>(pv, w);
    else
        k_enumerate_logdomain<float><<<grid, LD_THREADS, pool_bytes<float>(), s>>>(pv, w);
}

// --------------------------------------------------------------------- explicit items
template <typename T>
__global__ void __launch_bounds__(LD_THREADS) k_score_items_logdomain(ProblemView pv, const ScoreItem* __restrict__ items,
                                                                      double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* pool = reinterpret_cast<T*>(smem_raw);
    __shared__ T scratch[LD_THREADS / 32];
    __shared__ TreeProgram prog;
    __shared__ int s_ok;
    // grid = (items, samples): one CTA per (item, sample) so that a short list (the fp64 re-score of
    // the fp32 selection) still fills the GPU; the host adds the per-sample values in sample order
    const ScoreItem it = items[blockIdx.x];
    const int m = blockIdx.y;
    if (threadIdx.x == 0) s_ok = build_program(it.k, it.parent, it.nodes, &prog) ? 1 : 0;
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0) out[(size_t)blockIdx.x * pv.M + m] = CUDART_NAN;
        return;
    }
    const double score = eval_tree_logdomain<T>(pv, it.p, it.c, &prog, pool, scratch, m, m + 1);
    if (threadIdx.x == 0) out[(size_t)blockIdx.x * pv.M + m] = score;
}

void launch_score_items(int precision, const ProblemView& pv, const ScoreItem* d_items, int64_t n, double* d_out,
                        cudaStream_t s) {
    if (n <= 0) return;
    if (precision == BAMSE_FP64)
        k_score_items_logdomain<double><<<dim3((unsigned)n, pv.M), LD_THREADS, pool_bytes<double>(), s>>>(pv, d_items, d_out);
    else
        k_score_items_logdomain<float><<<dim3((unsigned)n, pv.M), LD_THREADS, pool_bytes<float>(), s>>>(pv, d_items, d_out);
}

// ------------------------------------------------------------------------ merge
// One CTA per parameter set over an unordered bag of records: topk rounds of a block-wide
// arg-best over the records of this parameter set that are strictly worse than the previous
// pick (the order total desc, clust asc, tree asc is strict, so no marks are needed).
__global__ void __launch_bounds__(512) k_merge_topk(const bamse_record* __restrict__ bag,
                                                    const unsigned long long* __restrict__ count,
                                                    unsigned long long n_given, int topk,
                                                    bamse_record* __restrict__ out) {
    constexpr int NT = 512;
    __shared__ bamse_record s_best[NT / 32];
    __shared__ bamse_record s_prev;
    __shared__ int s_has[NT / 32];
    const int p = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    // with a device-side count, n_given is the capacity of the bag (appends beyond it were dropped)
    const unsigned long long n = count ? (*count < n_given ? *count : n_given) : n_given;
    bool have_prev = false;
    bool done = false;
    for (int r = 0; r < topk; ++r) {
        bamse_record best; best.clust = -1; best.total = 0; best.score = 0; best.param = p; best.tree = 0;
        bool has = false;
        if (!done)
            for (unsigned long long i = t; i < n; i += NT) {
                const bamse_record c = bag[i];
                if (c.param != p || c.clust < 0 || c.total != c.total) continue;
                if (have_prev && !rec_better(s_prev, c)) continue;
                if (!has || rec_better(c, best)) { best = c; has = true; }
            }
        for (int d = 16; d > 0; d >>= 1) {
            bamse_record o;
            o.total = __shfl_xor_sync(0xffffffffu, best.total, d);
            o.score = __shfl_xor_sync(0xffffffffu, best.score, d);
            o.clust = __shfl_xor_sync(0xffffffffu, best.clust, d);
            o.param = __shfl_xor_sync(0xffffffffu, best.param, d);
            o.tree = __shfl_xor_sync(0xffffffffu, best.tree, d);
            const int oh = __shfl_xor_sync(0xffffffffu, (int)has, d);
            if (oh && (!has || rec_better(o, best))) { best = o; has = true; }
        }
        if (lane == 0) { s_best[w] = best; s_has[w] = has; }
        __syncthreads();
        if (t == 0) {
            bool h = false; bamse_record b = s_best[0];
            for (int i = 0; i < NT / 32; ++i)
                if (s_has[i] && (!h || rec_better(s_best[i], b))) { b = s_best[i]; h = true; }
            if (h) { s_prev = b; out[(size_t)p * topk + r] = b; }
            else {
                bamse_record e; e.total = -CUDART_INF; e.score = -CUDART_INF; e.clust = -1; e.param = p; e.tree = 0;
                out[(size_t)p * topk + r] = e;
            }
            s_has[0] = h;
        }
        __syncthreads();
        if (s_has[0]) have_prev = true; else done = true;
        __syncthreads();
    }
}

void launch_merge_topk(const bamse_record* d_bag, const unsigned long long* d_count, unsigned long long n, int P,
                       int topk, bamse_record* d_out, cudaStream_t s) {
    k_merge_topk<<<P, 512, 0, s>>>(d_bag, d_count, n, topk, d_out);
}

}  // namespace bamse
