// fp32 production kernels: table preparation and the warp-per-tree enumeration / list
// scoring kernels built on score_fast.cuh.
#include "kernels.cuh"
#include "score_fast.cuh"
#include "score_logdomain.cuh"

namespace bamse {

// ---------------------------------------------------------------- fast tables (fp64 -> fp32 log2)
// one WARP per (p, c, m, k): D2 = the binomial table as fp32 base-2 logs (rounded from the fp64 table), the leaf
// message EL = log2cumsum(D2) + p0 - log L and its scan SEL in rows k of the FG / FSG blocks (single-node forests),
// and the root table SD2 (suffix sums of d: the prefix scan of the reversed table), all with the production scan.
__global__ void __launch_bounds__(FAST_THREADS) k_tables_fast(ProblemView pv, FastTables ft, float* __restrict__ D2,
                                                              float* __restrict__ FG, float* __restrict__ FSG,
                                                              float* __restrict__ SD2) {
    __shared__ __align__(16) float vbuf[FAST_WARPS][L];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * FAST_WARPS + wid;
    if (row >= pv.P * pv.C * pv.M * pv.Kmax) return;
    const int k = row % pv.Kmax;
    const int m = (row / pv.Kmax) % pv.M;
    const int c = (row / (pv.Kmax * pv.M)) % pv.C;
    const int p = row / (pv.Kmax * pv.M * pv.C);
    if (k >= pv.K_of_c[c]) return;
    const TableBase tb = ft.tb[p * pv.C + c];
    if (!tb.mine) return;
    const size_t frow = (size_t)tb.fg_row + (size_t)m * tb.fi.nf + k;
    const size_t srow = (size_t)tb.fsg_row + (size_t)m * tb.nfsg + k;
    const float add2 = (float)((pv.p0[c] - 6.2383246250395077847550890931236) * 1.4426950408889634074);
    const double* src = pv.D64 + (size_t)row * L;
    float* v = vbuf[wid];
    float d[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        d[t] = (float)(src[lane + 32 * t] * 1.4426950408889634074);
        v[lane + 32 * t] = d[t];
        D2[(size_t)row * L + lane + 32 * t] = d[t];
    }
    __syncwarp();
    warp_log2cumsum(v, add2, lane);
    for (int t = 0; t < 4; ++t) reinterpret_cast<float4*>(FG + frow * L)[lane + 32 * t] = reinterpret_cast<const float4*>(v)[lane + 32 * t];
    warp_log2cumsum(v, add2, lane);
    for (int t = 0; t < 4; ++t) reinterpret_cast<float4*>(FSG + srow * L)[lane + 32 * t] = reinterpret_cast<const float4*>(v)[lane + 32 * t];
    // SD[i] = e^p0 / L^2 * sum_{x >= i} d[x]
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 16; ++t) v[L - 1 - lane - 32 * t] = d[t];
    __syncwarp();
    warp_log2cumsum(v, add2 - LOG2_L, lane);
#pragma unroll
    for (int t = 0; t < 16; ++t) SD2[(size_t)row * L + lane + 32 * t] = v[L - 1 - lane - 32 * t];
}

void launch_tables_fast(const ProblemView& pv, const FastTables& ft, float* D2, float* FG, float* FSG, float* SD2,
                        cudaStream_t s) {
    const int rows = pv.P * pv.C * pv.M * pv.Kmax;
    k_tables_fast<<<(rows + FAST_WARPS - 1) / FAST_WARPS, FAST_THREADS, 0, s>>>(pv, ft, D2, FG, FSG, SD2);
}

// Index -> program: executed by one lane once per tree.  Kept out of line so that its (large) code does not
// sit between the hot loops in the instruction cache.
static __device__ __noinline__ void build_tree_program(int K, uint64_t tree, FastProgram* prog, const ForestIndex* fi,
                                                       const int16_t* lut) {
    int8_t parent[KMAX];
    decode_tree(K, tree, parent);
    build_program_forest(K, parent, nullptr, *fi, lut, prog);
}

// geometry the table-driven programs of a clustering pack with (rows of forests with <= s_prog nodes are a
// prefix of the inside tables, so the row ids agree with the table geometry)
__device__ __forceinline__ ForestIndex program_index(const TableBase& tb) {
    ForestIndex f = tb.fi;
    f.s = tb.s_prog;
    if (f.sh > f.s) f.sh = f.s;
    return f;
}

// Program tables.  The program of a tree depends on (K, tree index) and on the table geometry (s, sh), not on
// the data, so for small K all K^(K-1) programs are built once, one THREAD per tree, and the
// enumeration warps fetch 64 bytes instead of running the (serial, one-lane) builder per tree.
__global__ void __launch_bounds__(128) k_build_prog_table(ForestIndex fi, uint64_t n_trees, const int16_t* __restrict__ lut,
                                                          FastProgram* __restrict__ out) {
    const uint64_t tree = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= n_trees) return;
    FastProgram prog;
    memset(&prog, 0, sizeof prog);
    build_tree_program(fi.K, tree, &prog, &fi, lut);
    out[tree] = prog;
}

void launch_build_prog_table(const ForestIndex& fi, uint64_t n_trees, const int16_t* d_lut, FastProgram* out, cudaStream_t s) {
    k_build_prog_table<<<(unsigned)((n_trees + 127) / 128), 128, 0, s>>>(fi, n_trees, d_lut, out);
}

// ------------------------------------------------------------------ per-warp evaluation
template <bool FRAMED, int SLOTS = FAST_SLOTS>
struct alignas(16) WarpSmemT {
    float vec[SLOTS][VSTRIDE];       // SLOTS = 2 suffices for K <= 6 (brute-forced), 3 for K <= 12
    uint16_t jstar[L];
    float fa[FRAMED ? FR_A : 4];      // framed-path scratch: 2^(a - frame) of the current segment
    float fb[FRAMED ? FR_SEG : 4];    // framed-path scratch: 2^(b - frame)
    FastProgram prog;
};
using WarpSmem = WarpSmemT<false>;
// dynamic shared memory: FAST_WARPS x WarpSmem
// the table builders hold two vectors (one convolution): 22.7 KB per CTA, 10 CTAs (40 warps) per SM
using BuildSmem = WarpSmemT<false, 2>;
#ifndef BAMSE_BUILD_CTAS
#define BAMSE_BUILD_CTAS 10
#endif

// Keeps the warp's shared-memory base in a register: without the opaque move ptxas
// rematerialises it from S2R tid / cta-id several times per convolution pass.
template <class W>
__device__ __forceinline__ W* pin_shared(W* p) {
    unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(s));
    return reinterpret_cast<W*>(__cvta_shared_to_generic(s));
}

template <class W>
__device__ inline void warp_init_guards(W* ws, int lane) {
    constexpr int n_slots = (int)(sizeof(ws->vec) / sizeof(ws->vec[0]));
    for (int s = 0; s < n_slots; ++s) {
        for (int i = lane; i < GUARD; i += 32) ws->vec[s][i] = -CUDART_INF_F;
        if (lane < 4) ws->vec[s][GUARD + L + lane] = -CUDART_INF_F;
    }
    __syncwarp();
}

// score of the tree whose program sits in ws->prog, summed over samples (fp64), natural log
// `cut`: early-exit bound on the sum over samples (-inf = evaluate everything).  Every per-sample
// score is <= 0 (a log of an average of products of probabilities <= 1), so once the partial sum is
// below the cut the tree cannot reach it; *n_done returns the samples actually evaluated.
template <bool FRAMED, class WS>
__device__ inline double warp_eval_tree(const ProblemView& pv, const FastTables& ft, int p, int c, WS* ws,
                                        int lane, unsigned* st_chunks, unsigned* st_convs, unsigned* st_scans,
                                        double cut = (double)-INFINITY, unsigned* n_done = nullptr, int m_begin = 0,
                                        int m_step = 1) {
    const float add2 = (float)((pv.p0[c] - 6.2383246250395077847550890931236) * 1.4426950408889634074);
    const int n_ops = ws->prog.n_ops;
    const TableBase* tb = ft.tb + (p * pv.C + c);
    const uint32_t fg_row = tb->fg_row, fh_row = tb->fh_row;
    const int nf = tb->fi.nf, nfh = tb->fi.nfh, Kc = tb->fi.K;
    double total = 0.0;
    unsigned done = 0;
    for (int m = m_begin; m < pv.M; m += m_step) {
        const size_t base = pv.table_offset(p, c, m);
        const size_t frow = (size_t)fg_row + (size_t)m * nf;
        const size_t srow = (size_t)tb->fsg_row + (size_t)m * tb->nfsg;
        ++done;
        int sp = 0;                                  // stack position == slot (convolution is in place)
        for (int i = 0; i < n_ops; ++i) {
            const int op = fast_op(ws->prog, i);
            const int arg = ws->prog.arg[i];
            if (op == FOP_PUSH_G || op == FOP_PUSH_SG) {
                BAMSE_BOUND(sp < (int)(sizeof(ws->vec) / sizeof(ws->vec[0])) && arg < (op == FOP_PUSH_G ? nf : tb->nfsg));
                float* dst = ws->vec[sp] + GUARD;
                warp_load_vec(dst, op == FOP_PUSH_G ? ft.FG + (frow + (size_t)arg) * L : ft.FSG + (srow + (size_t)arg) * L, lane);
                ++sp;
            } else if (op == FOP_SCAN) {
                warp_log2cumsum(ws->vec[sp - 1] + GUARD, add2, lane);
                ++*st_scans;
            } else if (op == FOP_CONV) {
                *st_chunks += warp_logconv<FRAMED>(ws->vec[sp - 2] + GUARD, ws->vec[sp - 1] + GUARD, ws->jstar, ws->fa, ws->fb, lane);
                ++*st_convs;
                --sp;
            } else {  // FOP_ADDD
                warp_add_table(ws->vec[sp - 1] + GUARD, ft.D2 + base + (size_t)arg * L, lane);
            }
        }
        const int fin = ws->prog.fin;
        const float* tab = nullptr;                  // FIN_LSE: plain log-sum-exp
        if (fin == FIN_DOT_D) tab = ft.D2 + base + (size_t)ws->prog.fin_r * L;
        else if (fin == FIN_DOT_SD) tab = ft.SD2 + base + (size_t)ws->prog.fin_r * L;
        BAMSE_BOUND(fin != FIN_DOT_FH || (int)ws->prog.fin_arg < nfh);
        BAMSE_BOUND(sp == 1);
        if (fin == FIN_DOT_FH)
            tab = ft.FH + ((size_t)fh_row + ((size_t)m * Kc + ws->prog.fin_r) * nfh + ws->prog.fin_arg) * L;
        float res2 = warp_log2dot(ws->vec[0] + GUARD, tab, lane);
        if (fin == FIN_LSE || fin == FIN_DOT_D) res2 -= LOG2_L;
        total += (double)res2 * 0.69314718055994530942;
        if (total < cut) {                           // warp-uniform: res2 is the same in every lane
            if (n_done) *n_done += done;
            return -CUDART_INF;
        }
    }
    if (n_done) *n_done += done;
    return total;
}

// Message tables, one level (= forests with n nodes) per launch: row <- rows with fewer nodes by its recipe.
// grid = (ceil(max rows of the level / FAST_WARPS), P * C * M); one warp per row, production convolution.
__global__ void __launch_bounds__(FAST_THREADS, BAMSE_BUILD_CTAS) k_forest_level(ProblemView pv, FastTables ft, int n, float* __restrict__ FG,
                                                               float* __restrict__ FSG, int fsg_top,
                                                               unsigned long long* stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = threadIdx.x >> 5;
    int lane = threadIdx.x & 31;
    BuildSmem* ws = pin_shared(reinterpret_cast<BuildSmem*>(smem_raw) + wid);
    // 1-D grid = (p, c, m) triples this rank builds tables for  x  row blocks (no 65 535 limit on either)
    const unsigned nblk = (unsigned)ft.grid_blocks, blk_x = blockIdx.x % nblk;
    const int pcm = ft.pcm_list[blockIdx.x / nblk];
    const int m = pcm % pv.M, pc = pcm / pv.M, c = pc % pv.C;
    const TableBase* tb = ft.tb + pc;
    if (n > tb->fi.s || !tb->mine) return;
    const int i = (int)(blk_x * FAST_WARPS + wid);
    if (i >= tb->fi.base[n + 1] - tb->fi.base[n]) return;
    // rows of a level in the order "convolutions first": the warps of a CTA get rows of the same kind
    const int row = (int)ft.level_order[tb->fi.K][tb->fi.base[n] + i];
    const ForestRecipe rec = ft.recipes[tb->fi.K][row];
    BAMSE_BOUND(row >= tb->fi.base[n] && row < tb->fi.base[n + 1] && (rec.kind == 1 ? (int)rec.b < tb->fi.base[n] && (int)rec.b < tb->nfsg
                                                                                      : (int)rec.a < tb->fi.base[n] && (int)rec.b < tb->fi.base[n]));
    // split (p, c): a top-level row is read only by the rank that scores the packs on its node set (when the
    // programs may push it too -- s_prog == s -- every rank keeps all rows)
    if (tb->split && n == tb->fi.s && tb->s_prog < n && set_owner(ft.pk + tb->fi.K, rec.mask, ft.shard_world) != ft.shard_rank) return;
    const size_t frow = (size_t)tb->fg_row + (size_t)m * tb->fi.nf;
    const size_t srow = (size_t)tb->fsg_row + (size_t)m * tb->nfsg;
    const float add2 = (float)((pv.p0[c] - 6.2383246250395077847550890931236) * 1.4426950408889634074);
    warp_init_guards(ws, lane);
    float* a = ws->vec[0] + GUARD;
    int chunks = 0, convs = 0;
    if (rec.kind == 1) {                            // one tree: E = D[u] + SG(children forest)
        warp_load_vec(a, FSG + (srow + rec.b) * L, lane);
        warp_add_table(a, ft.D2 + pv.table_offset(pc / pv.C, c, m) + (size_t)rec.u * L, lane);
    } else {                                        // several trees: the first one (*) the rest
        float* b = ws->vec[1] + GUARD;
        warp_load_vec(a, FG + (frow + rec.a) * L, lane);
        warp_load_vec(b, FG + (frow + rec.b) * L, lane);
        chunks = warp_logconv<false>(a, b, ws->jstar, ws->fa, ws->fb, lane);
        convs = 1;
    }
    float4* og = reinterpret_cast<float4*>(FG + (frow + row) * L);
    for (int t = 0; t < 4; ++t) og[lane + 32 * t] = reinterpret_cast<const float4*>(a)[lane + 32 * t];
    // the scanned message of a top-level forest is only read by the table-driven programs (fsg_top = 1)
    const bool scan = !(n == tb->fi.s && !fsg_top) && row < tb->nfsg;
    if (stats && lane == 0) {
        atomicAdd(stats + 0, 128ull * chunks + 512ull * convs + (scan ? 59ull * 32 : 0ull));
        atomicAdd(stats + 1, 128ull * chunks);
        atomicAdd(stats + 2, (unsigned long long)convs);
        atomicAdd(stats + 3, scan ? 1ull : 0ull);
    }
    if (!scan) return;
    warp_log2cumsum(a, add2, lane);
    float4* os = reinterpret_cast<float4*>(FSG + (srow + row) * L);
    for (int t = 0; t < 4; ++t) os[lane + 32 * t] = reinterpret_cast<const float4*>(a)[lane + 32 * t];
}

// Finishing tables FH[r][F][i] = 1/L^2 * sum_j d_r[i+j] * sg_F[j]: the correlation of D[r] with the scanned
// message of F is the convolution of the REVERSED (still log-concave) D[r] with it, read backwards.
// grid = (ceil(max K * nfh / FAST_WARPS), P * C * M); one warp per (root, forest) with the production convolution.
__global__ void __launch_bounds__(FAST_THREADS, BAMSE_BUILD_CTAS) k_forest_fh(ProblemView pv, FastTables ft, float* __restrict__ FH) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = threadIdx.x >> 5;
    int lane = threadIdx.x & 31;
    BuildSmem* ws = pin_shared(reinterpret_cast<BuildSmem*>(smem_raw) + wid);
    // 1-D grid = (p, c, m) triples this rank builds tables for  x  row blocks (no 65 535 limit on either)
    const unsigned nblk = (unsigned)ft.grid_blocks, blk_x = blockIdx.x % nblk;
    const int pcm = ft.pcm_list[blockIdx.x / nblk];
    const int m = pcm % pv.M, pc = pcm / pv.M, c = pc % pv.C;
    const TableBase* tb = ft.tb + pc;
    const int K = tb->fi.K, nfh = tb->fi.nfh;
    const int i = (int)(blk_x * FAST_WARPS + wid);
    if (i >= K * nfh || !tb->mine) return;
    const int r = i / nfh, f = i - r * nfh;
    if ((ft.recipes[K][f].mask >> r) & 1) return;   // the root is not part of its own children
    warp_init_guards(ws, lane);
    float* a = ws->vec[0] + GUARD;
    float* b = ws->vec[1] + GUARD;
    const float* D = ft.D2 + pv.table_offset(pc / pv.C, c, m) + (size_t)r * L;
    for (int t = 0; t < 16; ++t) a[lane + 32 * t] = __ldg(D + (L - 1 - lane - 32 * t));
    warp_load_vec(b, ft.FSG + ((size_t)tb->fsg_row + (size_t)m * tb->nfsg + f) * L, lane);
    warp_logconv<false>(a, b, ws->jstar, ws->fa, ws->fb, lane);
    float* out = FH + ((size_t)tb->fh_row + ((size_t)m * K + r) * nfh + f) * L;
    for (int t = 0; t < 16; ++t) out[lane + 32 * t] = a[L - 1 - lane - 32 * t] - LOG2_L;
}

// Scanned messages of the top-level forests, built on demand (the table-driven programs read them; the pair
// path does not).  grid = (ceil(max rows of level s / FAST_WARPS), P * C * M); one warp per row.
__global__ void __launch_bounds__(FAST_THREADS) k_forest_scan_top(ProblemView pv, FastTables ft, const float* __restrict__ FG,
                                                                  float* __restrict__ FSG) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = threadIdx.x >> 5;
    int lane = threadIdx.x & 31;
    BuildSmem* ws = pin_shared(reinterpret_cast<BuildSmem*>(smem_raw) + wid);
    // 1-D grid = (p, c, m) triples this rank builds tables for  x  row blocks (no 65 535 limit on either)
    const unsigned nblk = (unsigned)ft.grid_blocks, blk_x = blockIdx.x % nblk;
    const int pcm = ft.pcm_list[blockIdx.x / nblk];
    const int m = pcm % pv.M, pc = pcm / pv.M, c = pc % pv.C;
    const TableBase* tb = ft.tb + pc;
    const int n = tb->fi.s;
    if (tb->s_prog < n || !tb->mine) return;        // the programs of this clustering never touch the top level
    const int i = (int)(blk_x * FAST_WARPS + wid);
    if (i >= tb->fi.base[n + 1] - tb->fi.base[n]) return;
    const size_t row = (size_t)tb->fg_row + (size_t)m * tb->fi.nf + tb->fi.base[n] + i;
    const size_t srow = (size_t)tb->fsg_row + (size_t)m * tb->nfsg + tb->fi.base[n] + i;
    const float add2 = (float)((pv.p0[c] - 6.2383246250395077847550890931236) * 1.4426950408889634074);
    float* a = ws->vec[0] + GUARD;
    warp_load_vec(a, FG + row * L, lane);
    warp_log2cumsum(a, add2, lane);
    float4* os = reinterpret_cast<float4*>(FSG + srow * L);
    for (int t = 0; t < 4; ++t) os[lane + 32 * t] = reinterpret_cast<const float4*>(a)[lane + 32 * t];
}

// Context tables (pairs.cuh), one level (= contexts with q nodes) per launch: with O the outside message at the
// marked node v (the row of the parent context, 1/L at the root), A[y] = sum_{x >= y} O[x] d_v[x] and H the
// forest of v's children inside the context,  T[z] = e^p0 / L^2 sum_j g_H[j] A[z + j]  (no H: e^p0 / L A[z]).
// The suffix sum is the prefix scan of the reversed vector and the correlation the convolution of the reversed
// (still log-concave) A with g_H, read backwards.  grid = (ceil(max rows / FAST_WARPS), P * C * M); one warp per row.
__global__ void __launch_bounds__(FAST_THREADS, BAMSE_BUILD_CTAS) k_ctx_level(ProblemView pv, FastTables ft, int q, float* __restrict__ T,
                                                            unsigned long long* stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = threadIdx.x >> 5;
    int lane = threadIdx.x & 31;
    BuildSmem* ws = pin_shared(reinterpret_cast<BuildSmem*>(smem_raw) + wid);
    // 1-D grid = (p, c, m) triples this rank builds tables for  x  row blocks (no 65 535 limit on either)
    const unsigned nblk = (unsigned)ft.grid_blocks, blk_x = blockIdx.x % nblk;
    const int pcm = ft.pcm_list[blockIdx.x / nblk];
    const int m = pcm % pv.M, pc = pcm / pv.M, c = pc % pv.C;
    const TableBase* tb = ft.tb + pc;
    const PairK* pk = ft.pk + tb->fi.K;
    if (q > pk->g.r || !tb->mine) return;
    const uint32_t i = blk_x * FAST_WARPS + wid;
    if (i >= pk->lbase[q + 1] - pk->lbase[q]) return;
    const uint32_t row = pk->lbase[q] + i;
    const CtxRecipe rec = pk->recipes[row];
    BAMSE_BOUND(row < pk->nt && rec.parent_row < (int32_t)pk->lbase[q] && rec.h_row < tb->fi.nf && rec.v >= 0 && rec.v < tb->fi.K);
    // split (p, c): a top-level context is read only by the rank that scores the packs on the complement
    if (tb->split && q == pk->g.r &&
        set_owner(pk, ((1u << tb->fi.K) - 1u) & ~(uint32_t)rec.mask, ft.shard_world) != ft.shard_rank) return;
    const size_t trow = (size_t)tb->t_row + (size_t)m * pk->nt;
    const float add2 = (float)((pv.p0[c] - 6.2383246250395077847550890931236) * 1.4426950408889634074);
    warp_init_guards(ws, lane);
    float* a = ws->vec[0] + GUARD;
    const float* D = ft.D2 + pv.table_offset(pc / pv.C, c, m) + (size_t)rec.v * L;
    if (rec.parent_row >= 0) {
        const float* O = T + (trow + (size_t)rec.parent_row) * L;
        for (int t = 0; t < 16; ++t) a[lane + 32 * t] = __ldg(O + (L - 1 - lane - 32 * t)) + __ldg(D + (L - 1 - lane - 32 * t));
    } else {
        for (int t = 0; t < 16; ++t) a[lane + 32 * t] = __ldg(D + (L - 1 - lane - 32 * t)) - LOG2_L;
    }
    __syncwarp();
    warp_log2cumsum(a, 0.f, lane);
    if (rec.h_row >= 0) {
        float* b = ws->vec[1] + GUARD;
        warp_load_vec(b, ft.FG + ((size_t)tb->fg_row + (size_t)m * tb->fi.nf + (size_t)rec.h_row) * L, lane);
        const int chunks = warp_logconv<false>(a, b, ws->jstar, ws->fa, ws->fb, lane);
        if (stats && lane == 0) {
            atomicAdd(stats + 0, 128ull * chunks + 512ull);
            atomicAdd(stats + 1, 128ull * chunks);
            atomicAdd(stats + 2, 1ull);
        }
    }
    if (stats && lane == 0) { atomicAdd(stats + 0, 59ull * 32); atomicAdd(stats + 3, 1ull); }
    float* out = T + (trow + row) * L;
    for (int t = 0; t < 16; ++t) out[lane + 32 * t] = a[L - 1 - lane - 32 * t] + add2;
}

// Pair table of a cluster count (data independent), two passes over all tree indices: pass 1 (out == NULL) marks
// the context rows some regular tree refers to and counts the irregular trees; pass 2 (map != NULL) writes the
// entries with the compact context row.
__global__ void __launch_bounds__(128) k_build_pair_table(ForestIndex fi, PairGeom g, uint64_t n_trees,
                                                          const int16_t* __restrict__ lut, uint64_t* __restrict__ out,
                                                          uint8_t* __restrict__ used, unsigned long long* n_irregular,
                                                          const uint32_t* __restrict__ map, ForestIndex fi_prog,
                                                          FastProgram* __restrict__ irr_progs,
                                                          unsigned long long* __restrict__ mask_hist) {
    const uint64_t tree = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= n_trees) return;
    int8_t parent[KMAX];
    decode_tree(fi.K, tree, parent);
    uint32_t fg = 0, t = 0, pack = 0;
    const bool reg = classify_tree(fi.K, parent, fi, lut, g, &fg, &t, &pack);
    if (!out) {
        if (reg) used[t] = 1;
        else atomicAdd(n_irregular, 1ull);
        return;
    }
    uint64_t e = 0;
    if (reg) {
        TreeProgram ref;
        build_program(fi.K, parent, nullptr, &ref);
        e = pair_entry(fg, map[t], (uint32_t)ref.scre);
        if (mask_hist) atomicAdd(mask_hist + pack, 1ull);
    } else if (irr_progs) {
        // an irregular tree: its table-driven program, once and for all; the entry (scre field 0) carries its slot
        const unsigned long long slot = atomicAdd(n_irregular, 1ull);
        FastProgram prog;
        memset(&prog, 0, sizeof prog);
        build_program_forest(fi.K, parent, nullptr, fi_prog, lut, &prog);
        irr_progs[slot] = prog;
        e = (uint64_t)slot;
    }
    out[tree] = e;
}

// counting sort of a pair table by inside row: histogram, (host) prefix, scatter
__global__ void __launch_bounds__(256) k_pair_hist(const uint64_t* __restrict__ tab, uint64_t n_trees, unsigned* __restrict__ hist) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_trees) return;
    uint32_t fg, t, scre;
    pair_unpack(tab[i], &fg, &t, &scre);
    if (!scre) fg = 0;                                // an irregular tree (its entry holds a program slot): bin 0
    atomicAdd(hist + fg, 1u);
}

__global__ void __launch_bounds__(256) k_pair_scatter(const uint64_t* __restrict__ tab, uint64_t n_trees, unsigned* __restrict__ cursor,
                                                      uint32_t* __restrict__ sorted_tree, uint64_t* __restrict__ sorted_entry) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_trees) return;
    const uint64_t e = tab[i];
    uint32_t fg, t, scre;
    pair_unpack(e, &fg, &t, &scre);
    if (!scre) fg = 0;
    const unsigned pos = atomicAdd(cursor + fg, 1u);
    sorted_tree[pos] = (uint32_t)i;
    sorted_entry[pos] = e;
}

void launch_pair_hist(const uint64_t* tab, uint64_t n_trees, unsigned* hist, cudaStream_t s) {
    k_pair_hist<<<(unsigned)((n_trees + 255) / 256), 256, 0, s>>>(tab, n_trees, hist);
}

void launch_pair_scatter(const uint64_t* tab, uint64_t n_trees, unsigned* cursor, uint32_t* sorted_tree, uint64_t* sorted_entry,
                         cudaStream_t s) {
    k_pair_scatter<<<(unsigned)((n_trees + 255) / 256), 256, 0, s>>>(tab, n_trees, cursor, sorted_tree, sorted_entry);
}

__global__ void k_flat_prefix(int P, const unsigned long long* __restrict__ cnt, const unsigned long long* __restrict__ off,
                              unsigned long long* __restrict__ prefix) {
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long acc = 0;
    for (int p = 0; p < P; ++p) {
        prefix[p] = acc;
        const unsigned long long cap = off[p + 1] - off[p];
        acc += cnt[p] < cap ? cnt[p] : cap;             // an overflowing region is reported through flat_overflow
    }
    prefix[P] = acc;
}

// a record written by another warp of the CTA (global memory, read after the barrier): bypass L1
__device__ __forceinline__ bamse_record load_record_cg(const bamse_record* p) {
    static_assert(sizeof(bamse_record) == 32, "two 16-byte loads");
    union { int4 q[2]; bamse_record r; } u;
    u.q[0] = __ldcg(reinterpret_cast<const int4*>(p));
    u.q[1] = __ldcg(reinterpret_cast<const int4*>(p) + 1);
    return u.r;
}

__device__ __forceinline__ bool rec_better_f(const bamse_record& a, const bamse_record& b) {
    if (a.total != b.total) return a.total > b.total;
    if (a.clust != b.clust) return a.clust < b.clust;
    return a.tree < b.tree;
}

// ---------------------------------------------------------------------- enumerate (fp32)
// Persistent warps: every warp pulls small chunks of consecutive tree indices from one global
// counter and scores them on its own -- no block-level barrier anywhere, so a warp that drew
// cheap trees (few leaves) simply pulls more, and the tail of the launch is one tree long.
// Each warp keeps a private top-k list (global memory, EnumWork::lists) and appends it to a flat global bag
// when its parameter set changes and at the end; k_merge_topk reduces the bag.
#ifndef BAMSE_CTAS_2SLOT
#define BAMSE_CTAS_2SLOT 10     // resident CTAs per SM of the 2-slot instantiation (48 registers, 40 warps, 22.7 KB each)
#endif
template <bool FRAMED, int SLOTS>
__global__ void __launch_bounds__(FAST_THREADS, FRAMED ? 6 : (SLOTS == 2 ? BAMSE_CTAS_2SLOT : 7))
k_enumerate_fast(ProblemView pv, FastTables ft, EnumWork w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using WS = WarpSmemT<FRAMED, SLOTS>;
    WS* wsm = reinterpret_cast<WS*>(smem_raw);
    const int t = threadIdx.x, wid = t >> 5;
    int lane = t & 31;
    asm volatile("" : "+r"(lane));
    WS* ws = pin_shared(wsm + wid);
    const int kcap = w.topk > 0 ? w.topk : 1;
    bamse_record* const cta_lists = w.lists + (size_t)blockIdx.x * FAST_WARPS * kcap;
    bamse_record* list = cta_lists + wid * kcap;
    warp_init_guards(ws, lane);
    int ntop = 0;                                    // meaningful in lane 0
    int cur_p = -1;
    // flat mode: the work items are the entries of the per-parameter-set lists k_enumerate_pairs left behind
    // (irregular trees), one tree per chunk
    const bool flat = w.flat_mode != 0;
    const uint64_t total = flat ? w.flat_prefix[pv.P] : *w.total_chunks;
    unsigned st_chunks = 0, st_convs = 0, st_scans = 0, st_samples = 0;

    auto flush = [&]() {
        if (lane == 0 && ntop > 0) {
            const unsigned long long pos = atomicAdd(w.cand_count, (unsigned long long)ntop);
            if (pos + ntop <= w.cand_cap)
                for (int i = 0; i < ntop; ++i) w.cand[pos + i] = list[i];
            ntop = 0;
        }
    };

    // Tail splitting: the last n_split single-tree chunks are issued as split_ways items each (every
    // item takes the samples m = h, h + ways, ...); the parts meet in split_acc / split_cnt and the part
    // that finishes last ranks the tree.  With few trees per warp (multi-GPU shards of a small
    // problem) this shortens the tail of the launch from one tree to a fraction of one.
    const int ways = w.split_ways > 1 ? w.split_ways : 1;
    const uint64_t n_split = ways > 1 ? (w.n_split < total ? w.n_split : total) : 0;
    const uint64_t split_from = total - n_split;
    const uint64_t vtotal = split_from + n_split * (uint64_t)ways;
    for (;;) {
        unsigned long long chunk = 0;
        if (lane == 0) chunk = atomicAdd(w.next_chunk, 1ull);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= vtotal) break;
        int part = 0, step = 1;
        uint64_t split_slot = 0;
        if (chunk >= split_from) {
            const uint64_t k = chunk - split_from;
            split_slot = k / (uint64_t)ways;
            part = (int)(k - split_slot * (uint64_t)ways);
            step = ways;
            chunk = split_from + split_slot;
        }
        int lo = 0, hi = (flat ? pv.P : w.n_segs) - 1;
        uint64_t flat_tree = 0;
        if (flat) {
            while (lo < hi) {                        // parameter set of this entry
                const int mid = (lo + hi + 1) >> 1;
                if (w.flat_prefix[mid] <= chunk) lo = mid; else hi = mid - 1;
            }
            BAMSE_BOUND(w.flat_off[lo] + (chunk - w.flat_prefix[lo]) < w.flat_off[lo + 1]);
            const unsigned long long ent = w.flat[w.flat_off[lo] + (chunk - w.flat_prefix[lo])];
            lo = (int)(ent >> 40);
            BAMSE_BOUND(lo < w.n_segs);
            flat_tree = ent & ((1ull << 40) - 1ull);
        } else {
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (w.segs[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
            }
        }
        const Segment seg = w.segs[lo];
        if (seg.p != cur_p) { flush(); cur_p = seg.p; }
        const uint64_t off = flat ? (flat_tree - seg.tree_begin) / (uint64_t)seg.stride
                                  : (chunk - seg.first_chunk) * (uint64_t)w.chunk;
        const uint64_t remain = seg.n_trees - off;
        const int n = flat ? 1 : (remain < (uint64_t)w.chunk ? (int)remain : w.chunk);
        const FastProgram* tab = w.prog_tabs ? w.prog_tabs[seg.K] : nullptr;
        const PairK* pkf = ft.pk + seg.K;
        for (int i = 0; i < n; ++i) {
            const uint64_t tree = seg.tree_begin + (off + i) * (uint64_t)seg.stride;
            const FastProgram* src = tab ? tab + tree : nullptr;
            if (flat && pkf->irr_progs && pkf->pair_tab)         // the list holds irregular trees only: their slot is in the entry
                src = pkf->irr_progs + (__ldg(pkf->pair_tab + tree) & ((1ull << (PAIR_FG_BITS + PAIR_T_BITS)) - 1ull));
            if (src) {
                if (lane < 16)
                    reinterpret_cast<uint32_t*>(&ws->prog)[lane] = __ldg(reinterpret_cast<const uint32_t*>(src) + lane);
            } else if (lane == 0) {
                const ForestIndex fip = program_index(ft.tb[seg.p * pv.C + seg.c]);
                build_tree_program(seg.K, tree, &ws->prog, &fip, ft.lut);
            }
            __syncwarp();
            if (ws->prog.max_sp > SLOTS) asm volatile("trap;");     // a program this instantiation cannot hold: fail loudly
            double cut = -CUDART_INF;
            if (w.early_exit) {
                // lane 0 knows the warp's k-th best and this tree's log scre; fp32 slack of 1e-3 nats
                if (lane == 0) {
                    double c0 = seg.threshold;
                    if (w.topk > 0 && ntop == w.topk) c0 = fmax(c0, list[w.topk - 1].total - seg.clust_const);
                    cut = c0 - log((double)ws->prog.scre) - 1e-3 - 1e-6 * fabs(c0);
                }
                cut = __shfl_sync(0xffffffffu, cut, 0);
            }
            double score = warp_eval_tree<FRAMED, WS>(pv, ft, seg.p, seg.c, ws, lane, &st_chunks, &st_convs, &st_scans,
                                                      cut, &st_samples, part, step);
            if (step > 1) {
                // combine the parts; only the last one to arrive goes on (two or four addends per
                // tree: the sum does not depend on the arrival order beyond the last bit)
                int last = 0;
                if (lane == 0) {
                    atomicAdd(w.split_acc + split_slot, score);
                    __threadfence();
                    last = atomicAdd(w.split_cnt + split_slot, 1) == step - 1;
                    if (last) score = atomicAdd(w.split_acc + split_slot, 0.0);
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (!last) { __syncwarp(); continue; }
            }
            if (lane == 0) {
                const double logscre = log((double)ws->prog.scre);
                const double full = score + logscre;
                if (w.dense_score) w.dense_score[off + i] = score;
                if (w.dense_logscre) w.dense_logscre[off + i] = logscre;
                if (w.topk > 0 && full >= seg.threshold && full > -CUDART_INF) {   // an early-exited tree is -inf
                    bamse_record r;
                    r.total = full + seg.clust_const; r.score = full; r.clust = seg.c; r.param = seg.p; r.tree = tree;
                    if (!(ntop == w.topk && !rec_better_f(r, list[w.topk - 1]))) {
                        int q = ntop < w.topk ? ntop : w.topk - 1;
                        while (q > 0 && rec_better_f(r, list[q - 1])) { list[q] = list[q - 1]; --q; }
                        list[q] = r;
                        if (ntop < w.topk) ++ntop;
                    }
                }
            }
            __syncwarp();
        }
    }
    // end of the launch: the warps of a CTA finish within one tree of each other, so a block-level
    // merge of their lists is cheap and keeps the bag (and the merge kernel's input) 4x smaller
    // (only records of the same parameter set compete; with mixed sets everything is appended)
    __shared__ int s_ntop[FAST_WARPS];
    __shared__ int s_p[FAST_WARPS];
    if (lane == 0) { s_ntop[wid] = ntop; s_p[wid] = cur_p; }
    __syncthreads();
    if (t == 0 && w.topk > 0) {
        int pos[FAST_WARPS];
        int total_n = 0, p_seen = -1;
        bool same_p = true;
        for (int i = 0; i < FAST_WARPS; ++i) {
            pos[i] = 0; total_n += s_ntop[i];
            if (s_ntop[i] > 0) { if (p_seen < 0) p_seen = s_p[i]; else if (s_p[i] != p_seen) same_p = false; }
        }
        const int n_out = (same_p && total_n > w.topk) ? w.topk : total_n;
        if (n_out > 0) {
            const unsigned long long base = atomicAdd(w.cand_count, (unsigned long long)n_out);
            for (int r = 0; r < n_out; ++r) {
                int best = -1;
                for (int i = 0; i < FAST_WARPS; ++i)
                    if (pos[i] < s_ntop[i] &&
                        (best < 0 || rec_better_f(load_record_cg(cta_lists + i * kcap + pos[i]),
                                                  load_record_cg(cta_lists + best * kcap + pos[best]))))
                        best = i;
                if (base + r < w.cand_cap) w.cand[base + r] = load_record_cg(cta_lists + best * kcap + pos[best]);
                ++pos[best];
            }
        }
    }
    if (w.stats && lane == 0) {
        // thread-level MUFU ops: 4 ex2 per chunk per lane x 32 lanes, 16 passes x 1 lg2 x 32 per conv,
        // 59 per lane per scan, 17 per lane per final log-sum-exp
        const unsigned long long mufu = 128ull * st_chunks + 512ull * st_convs + 59ull * 32 * st_scans + 17ull * 32 * st_samples;
        atomicAdd(w.stats + 0, mufu);
        atomicAdd(w.stats + 1, 128ull * st_chunks);
        atomicAdd(w.stats + 2, (unsigned long long)st_convs);
        atomicAdd(w.stats + 3, (unsigned long long)st_scans);
    }
    if (w.samples_done && lane == 0) atomicAdd(w.samples_done, (unsigned long long)st_samples);
}

// ---------------------------------------------------------------------- enumerate: pair path (fp32)
// log2 sum_i 2^(g[i] + t[i]) of two table rows in global memory, by ONE lane.  Both rows are log-concave, so
// the sum is unimodal: a binary search on the sign of its slope finds the maximising index, then aligned groups
// of four terms are added outwards on both sides until the group's term nearest to the maximum is below
// 2^-THETA of it (the terms decay at least geometrically from there: the same bound as the convolution windows).
__device__ __forceinline__ float lane_log2dot(const float* __restrict__ g, const float* __restrict__ t, unsigned& groups) {
    // search over the EVEN positions only: (v[2e], v[2e+1]) is one aligned 8-byte load per row, and the group of
    // four that holds the last even position with a non-negative slope (or its right neighbour) holds the maximum.
    // The kernel is bound by L1 wavefronts (every lane touches its own sector), so loads are what is counted.
    const float2* g2 = reinterpret_cast<const float2*>(g);
    const float2* t2 = reinterpret_cast<const float2*>(t);
    int lo = 0, hi = L / 2;                           // first pair whose slope is negative, in [0, 256]
#pragma unroll 1
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float2 a2 = __ldg(g2 + mid), b2 = __ldg(t2 + mid);
        if ((a2.y + b2.y) >= (a2.x + b2.x)) lo = mid + 1; else hi = mid;
    }
    // the maximum is at index 2 lo - 1, 2 lo or (slope between the pairs) just left of it: take the group of the
    // last rising pair and let the outward walk find the rest
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* t4 = reinterpret_cast<const float4*>(t);
    int q0 = lo > 0 ? (2 * lo - 1) >> 2 : 0;
    {   // the true maximum may sit in the next group (2 lo is its first element): compare the two group maxima
        if (2 * lo < L && ((2 * lo) >> 2) != q0) {
            const float2 c2 = __ldg(g2 + lo), d2 = __ldg(t2 + lo);
            const float2 e2 = __ldg(g2 + lo - 1), f2 = __ldg(t2 + lo - 1);
            if ((c2.x + d2.x) > (e2.y + f2.y)) q0 = (2 * lo) >> 2;
        }
    }
    float4 a = __ldg(g4 + q0), b = __ldg(t4 + q0);
    float x0 = a.x + b.x, x1 = a.y + b.y, x2 = a.z + b.z, x3 = a.w + b.w;
    const float m0 = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
    if (!(m0 > -CUDART_INF_F)) return -CUDART_INF_F;
    float sum = (ex2f(x0 - m0) + ex2f(x1 - m0)) + (ex2f(x2 - m0) + ex2f(x3 - m0));
    unsigned n = 1;
#pragma unroll 1
    for (int q = q0 + 1; q < L / 4; ++q) {
        a = __ldg(g4 + q); b = __ldg(t4 + q);
        x0 = (a.x + b.x) - m0;
        if (!(x0 >= -THETA)) break;
        x1 = (a.y + b.y) - m0; x2 = (a.z + b.z) - m0; x3 = (a.w + b.w) - m0;
        sum += (ex2f(x0) + ex2f(x1)) + (ex2f(x2) + ex2f(x3));
        ++n;
    }
#pragma unroll 1
    for (int q = q0 - 1; q >= 0; --q) {
        a = __ldg(g4 + q); b = __ldg(t4 + q);
        x3 = (a.w + b.w) - m0;
        if (!(x3 >= -THETA)) break;
        x0 = (a.x + b.x) - m0; x1 = (a.y + b.y) - m0; x2 = (a.z + b.z) - m0;
        sum += (ex2f(x0) + ex2f(x1)) + (ex2f(x2) + ex2f(x3));
        ++n;
    }
    groups += n;
    return m0 + lg2f(sum);
}

constexpr int PAIR_WARPS = 8;
constexpr int PAIR_THREADS = PAIR_WARPS * 32;

// One LANE per tree: persistent warps pull chunks of 32 consecutive tree slots of a segment.  A regular tree is
// two row ids (pair table, or the cut computed here for cluster counts without a table) and costs one windowed
// dot product per sample; an irregular tree is appended to the flat list of its parameter set for
// k_enumerate_fast.  Per-warp top-k lists (global memory) are filled by lane 0 from the lanes that beat the
// warp's k-th best, and appended to the candidate bag when the parameter set changes and at the end.
//
// MODE 0, tree-major: a lane takes its tree through all samples (early exit possible; the only mode for cluster
//   counts whose tables per sample exceed the L2 anyway).
// MODE 1 + MODE 2, sample-major: the chunks of a segment are ordered sample by sample, so the warps in flight read
//   the tables of ONE sample of one clustering (K = 7: 16 MB, L2 resident) instead of all samples at once (DRAM
//   traffic 2.5x the tables); MODE 1 stores the per-sample log2 results in `part` [M][slots], MODE 2 sums them in
//   sample order -- the same fp64 additions as MODE 0, bit for bit -- and ranks.
template <int MODE>
__global__ void __launch_bounds__(PAIR_THREADS, 6) k_enumerate_pairs(ProblemView pv, FastTables ft, EnumWork w) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * PAIR_WARPS + (threadIdx.x >> 5);
    const int kcap = w.topk > 0 ? w.topk : 1;
    bamse_record* list = w.lists2 + (size_t)gw * kcap;
    int ntop = 0;                                    // meaningful in lane 0
    int cur_p = -1;
    double kth = -CUDART_INF;                        // total of the warp's k-th best once the list is full (all lanes)
    const int reps = MODE == 0 ? 1 : pv.M;           // chunk counts of the segments are multiples of `reps`
    const uint64_t total_all = *w.total_chunks;
    const uint64_t total = MODE == 2 ? total_all / (uint64_t)reps : total_all;
    const uint64_t slots = total_all / (uint64_t)reps * 32ull;      // tree slots of the launch (sample-major)
    unsigned st_groups = 0, st_samples = 0, st_trees = 0;

    auto flush = [&]() {
        if (lane == 0 && ntop > 0) {
            const unsigned long long pos = atomicAdd(w.cand_count, (unsigned long long)ntop);
            if (pos + ntop <= w.cand_cap)
                for (int i = 0; i < ntop; ++i) w.cand[pos + i] = list[i];
            ntop = 0;
        }
    };

    // `grab` consecutive chunks per visit of the global counter: with 10^8 chunks (K = 10) one atomic per chunk on
    // one address costs ~1 ns each, on every rank -- a fixed cost that does not shrink with the number of GPUs
    const unsigned long long grab = w.grab > 1 ? (unsigned long long)w.grab : 1ull;
    unsigned long long chunk_base = 0, chunk_left = 0;
    for (;;) {
        if (chunk_left == 0) {
            if (lane == 0) chunk_base = atomicAdd(w.next_chunk, grab);
            chunk_base = __shfl_sync(0xffffffffu, chunk_base, 0);
            chunk_left = grab;
        }
        const unsigned long long chunk = chunk_base++;
        --chunk_left;
        if (chunk >= total) break;
        const uint64_t key = MODE == 2 ? chunk * (uint64_t)reps : chunk;
        int lo = 0, hi = w.n_segs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (w.segs[mid].first_chunk <= key) lo = mid; else hi = mid - 1;
        }
        const Segment seg = w.segs[lo];
        if (MODE != 1 && seg.p != cur_p) { flush(); cur_p = seg.p; kth = -CUDART_INF; }
        // position inside the segment: tree block and (sample-major accumulation) the sample
        uint64_t blk = key - seg.first_chunk;
        int m_one = 0;
        if (MODE == 1) {
            const uint64_t cpm = (seg.n_trees + 31ull) / 32ull;
            m_one = (int)(blk / cpm);
            blk -= (uint64_t)m_one * cpm;
        } else if (MODE == 2) blk /= (uint64_t)reps;
        const uint64_t off = blk * 32ull + (uint64_t)lane;
        const uint64_t slot = seg.first_chunk / (uint64_t)reps * 32ull + off;
        const bool valid = off < seg.n_trees;
        uint64_t tree = seg.tree_begin + off * (uint64_t)seg.stride;
        const TableBase* tb = ft.tb + (seg.p * pv.C + seg.c);
        const PairK* pk = ft.pk + seg.K;
        uint32_t fg = 0, tr = 0, scre = 0;
        bool reg = false;
        // a whole index range goes in the order of the inside rows (the dense debug outputs keep the index order)
        const bool sorted = pk->sorted_tree && seg.stride == 1 && seg.tree_begin == 0 && seg.whole && !w.dense_score;
        if (valid) {
            if (sorted) {
                tree = __ldg(pk->sorted_tree + off);
                pair_unpack(__ldg(pk->sorted_entry + off), &fg, &tr, &scre);
                reg = scre != 0;
            } else if (pk->pair_tab) {
                pair_unpack(__ldg(pk->pair_tab + tree), &fg, &tr, &scre);
                reg = scre != 0;
            } else if (pk->g.r > 0) {
                int8_t parent[KMAX];
                decode_tree(seg.K, tree, parent);
                reg = classify_tree(seg.K, parent, tb->fi, ft.lut, pk->g, &fg, &tr);
                if (reg) {
                    TreeProgram ref;
                    build_program(seg.K, parent, nullptr, &ref);
                    scre = (uint32_t)ref.scre;
                    tr = __ldg(pk->map + tr);
                }
            }
        }
        // a (p, c) split over all ranks: every rank walks all tree indices (8 bytes each), takes the regular trees
        // whose pack sits on one of its node sets and every world-th irregular tree
        bool mine = true;
        if (tb->split && valid) {
            if (reg) mine = set_owner(pk, ft.recipes[seg.K][fg].mask, ft.shard_world) == ft.shard_rank;
            else mine = (int)(tree % (uint64_t)ft.shard_world) == ft.shard_rank;
        }
        reg = reg && mine;
        if (MODE != 1 || m_one == 0) st_trees += (valid && mine) ? 1u : 0u;
        // irregular trees: to the flat list of this parameter set (warp-aggregated append), once
        if (MODE == 0 || (MODE == 1 && m_one == 0)) {
            const bool irr = valid && mine && !reg;
            const unsigned irr_mask = __ballot_sync(0xffffffffu, irr);
            if (irr_mask) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(w.flat_cnt + seg.p, (unsigned long long)__popc(irr_mask));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (irr) {
                    const unsigned long long pos = w.flat_off[seg.p] + base + (unsigned long long)__popc(irr_mask & ((1u << lane) - 1u));
                    if (pos < w.flat_off[seg.p + 1]) w.flat[pos] = ((unsigned long long)lo << 40) | tree;
                    else *w.flat_overflow = 1ull;
                }
            }
        }
        const float* g = ft.FG + ((size_t)tb->fg_row + fg) * L;
        const float* t = ft.T + ((size_t)tb->t_row + tr) * L;
        const size_t gstep = (size_t)tb->fi.nf * L, tstep = (size_t)pk->nt * L;
        BAMSE_BOUND(!reg || ((int)fg < tb->fi.nf && tr < pk->nt && tb->mine));
        BAMSE_BOUND(MODE == 0 || !reg || slot < slots);
        if (MODE == 1) {
            if (reg) {
                w.part[(size_t)m_one * slots + slot] = lane_log2dot(g + (size_t)m_one * gstep, t + (size_t)m_one * tstep, st_groups);
                ++st_samples;
            }
            continue;
        }
        double score = -CUDART_INF, logscre = 0.0;
        if (reg) {
            logscre = log((double)scre);
            double acc = 0.0;
            if (MODE == 0) {
                double cut = -CUDART_INF;
                if (w.early_exit) {
                    const double c0 = fmax(seg.threshold, kth - seg.clust_const);
                    cut = c0 - logscre - 1e-3 - 1e-6 * fabs(c0);
                }
#pragma unroll 1
                for (int m = 0; m < pv.M; ++m) {
                    acc += (double)lane_log2dot(g, t, st_groups) * 0.69314718055994530942;
                    ++st_samples;
                    g += gstep; t += tstep;
                    if (acc < cut) { acc = -CUDART_INF; break; }
                }
            } else {
#pragma unroll 1
                for (int m = 0; m < pv.M; ++m) acc += (double)w.part[(size_t)m * slots + slot] * 0.69314718055994530942;
            }
            score = acc;
            if (w.dense_score) w.dense_score[off] = score;
            if (w.dense_logscre) w.dense_logscre[off] = logscre;
        }
        if (w.topk > 0) {
            const double full = score + logscre;
            const double tot = full + seg.clust_const;
            unsigned cmask = __ballot_sync(0xffffffffu, reg && full >= seg.threshold && full > -CUDART_INF && tot >= kth);
            while (cmask) {
                const int src = __ffs(cmask) - 1;
                cmask &= cmask - 1;
                const double r_full = __shfl_sync(0xffffffffu, full, src);
                const unsigned long long r_tree = __shfl_sync(0xffffffffu, (unsigned long long)tree, src);
                if (lane == 0) {
                    bamse_record r;
                    r.total = r_full + seg.clust_const; r.score = r_full; r.clust = seg.c; r.param = seg.p; r.tree = r_tree;
                    if (!(ntop == w.topk && !rec_better_f(r, list[w.topk - 1]))) {
                        int q = ntop < w.topk ? ntop : w.topk - 1;
                        while (q > 0 && rec_better_f(r, list[q - 1])) { list[q] = list[q - 1]; --q; }
                        list[q] = r;
                        if (ntop < w.topk) ++ntop;
                    }
                    if (ntop == w.topk) kth = list[w.topk - 1].total;
                }
                kth = __shfl_sync(0xffffffffu, kth, 0);
            }
        }
    }
    if (MODE != 1) flush();
    // executed-work counters: 4 ex2 per group, 1 lg2 per dot
    st_groups = __reduce_add_sync(0xffffffffu, st_groups);
    st_samples = __reduce_add_sync(0xffffffffu, st_samples);
    if (MODE != 2) {
        if (w.stats && lane == 0) atomicAdd(w.stats + 0, 4ull * st_groups + st_samples);
        if (w.samples_done && lane == 0) atomicAdd(w.samples_done, (unsigned long long)st_samples);
        st_trees = __reduce_add_sync(0xffffffffu, st_trees);
        if (w.trees_done && lane == 0 && st_trees) atomicAdd(w.trees_done, (unsigned long long)st_trees);
    }
}

int enumerate_pairs_grid_size(int device) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_enumerate_pairs<0>, PAIR_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

int enumerate_pairs_workers(int grid) { return grid * PAIR_WARPS; }

void launch_enumerate_pairs(const ProblemView& pv, const FastTables& ft, const EnumWork& w, int grid, int mode, cudaStream_t s) {
    if (mode == 0) k_enumerate_pairs<0><<<grid, PAIR_THREADS, 0, s>>>(pv, ft, w);
    else if (mode == 1) k_enumerate_pairs<1><<<grid, PAIR_THREADS, 0, s>>>(pv, ft, w);
    else k_enumerate_pairs<2><<<grid, PAIR_THREADS, 0, s>>>(pv, ft, w);
}

void launch_flat_prefix(int P, const unsigned long long* cnt, const unsigned long long* off, unsigned long long* prefix,
                        cudaStream_t s) {
    k_flat_prefix<<<1, 32, 0, s>>>(P, cnt, off, prefix);
}

void launch_ctx_level(const ProblemView& pv, const FastTables& ft, int q, int max_rows, float* T, unsigned long long* stats,
                      cudaStream_t s) {
    if (max_rows <= 0) return;
    const size_t smem = sizeof(BuildSmem) * FAST_WARPS;
    cudaFuncSetAttribute(k_ctx_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FastTables ftg = ft;
    ftg.grid_blocks = (max_rows + FAST_WARPS - 1) / FAST_WARPS;
    const unsigned grid = (unsigned)ftg.grid_blocks * (unsigned)ft.n_pcm;
    k_ctx_level<<<grid, FAST_THREADS, smem, s>>>(pv, ftg, q, T, stats);
}

void launch_forest_scan_top(const ProblemView& pv, const FastTables& ft, int max_rows, const float* FG, float* FSG,
                            cudaStream_t s) {
    if (max_rows <= 0) return;
    const size_t smem = sizeof(BuildSmem) * FAST_WARPS;
    cudaFuncSetAttribute(k_forest_scan_top, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FastTables ftg = ft;
    ftg.grid_blocks = (max_rows + FAST_WARPS - 1) / FAST_WARPS;
    const unsigned grid = (unsigned)ftg.grid_blocks * (unsigned)ft.n_pcm;
    k_forest_scan_top<<<grid, FAST_THREADS, smem, s>>>(pv, ftg, FG, FSG);
}

void launch_build_pair_table(const ForestIndex& fi, const PairGeom& g, uint64_t n_trees, const int16_t* d_lut,
                             uint64_t* out, uint8_t* used, unsigned long long* n_irregular, const uint32_t* d_map,
                             const ForestIndex& fi_prog, FastProgram* irr_progs, unsigned long long* mask_hist, cudaStream_t s) {
    k_build_pair_table<<<(unsigned)((n_trees + 127) / 128), 128, 0, s>>>(fi, g, n_trees, d_lut, out, used, n_irregular, d_map,
                                                                         fi_prog, irr_progs, mask_hist);
}

static size_t fast_smem_bytes(int topk, bool framed = false, int slots = FAST_SLOTS) {
    const size_t per_warp = framed ? sizeof(WarpSmemT<true, 3>) : (slots == 2 ? sizeof(WarpSmemT<false, 2>) : sizeof(WarpSmemT<false, 3>));
    (void)topk;                                      // the top-k lists live in global memory (EnumWork::lists)
    return per_warp * FAST_WARPS;
}

void launch_forest_level(const ProblemView& pv, const FastTables& ft, int n, int max_rows, float* FG, float* FSG,
                         int fsg_top, unsigned long long* stats, cudaStream_t s) {
    if (max_rows <= 0) return;
    const size_t smem = sizeof(BuildSmem) * FAST_WARPS;
    cudaFuncSetAttribute(k_forest_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FastTables ftg = ft;
    ftg.grid_blocks = (max_rows + FAST_WARPS - 1) / FAST_WARPS;
    const unsigned grid = (unsigned)ftg.grid_blocks * (unsigned)ft.n_pcm;
    k_forest_level<<<grid, FAST_THREADS, smem, s>>>(pv, ftg, n, FG, FSG, fsg_top, stats);
}

void launch_forest_fh(const ProblemView& pv, const FastTables& ft, int max_items, float* FH, cudaStream_t s) {
    if (max_items <= 0) return;
    const size_t smem = sizeof(BuildSmem) * FAST_WARPS;
    cudaFuncSetAttribute(k_forest_fh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FastTables ftg = ft;
    ftg.grid_blocks = (max_items + FAST_WARPS - 1) / FAST_WARPS;
    const unsigned grid = (unsigned)ftg.grid_blocks * (unsigned)ft.n_pcm;
    k_forest_fh<<<grid, FAST_THREADS, smem, s>>>(pv, ftg, FH);
}

int enumerate_workers(int precision, int grid) { return precision == BAMSE_FP32 ? grid * FAST_WARPS : grid; }

// two vector slots suffice when the tables fold every small subtree (forest_slots_needed) -> 10 CTAs (40 warps)
// per SM instead of 7; the framed instantiation always carries three
static int fast_slots(int slots_needed, bool framed) { return (!framed && slots_needed <= 2) ? 2 : 3; }

int enumerate_fast_grid_size(int device, int topk, bool framed, int slots_needed) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int slots = fast_slots(slots_needed, framed);
    const size_t smax = fast_smem_bytes(BAMSE_TOPK_MAX, framed, slots), sm = fast_smem_bytes(topk, framed, slots);
    if (framed) {
        cudaFuncSetAttribute(k_enumerate_fast<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_enumerate_fast<true, 3>, FAST_THREADS, sm);
    } else if (slots == 2) {
        cudaFuncSetAttribute(k_enumerate_fast<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_enumerate_fast<false, 2>, FAST_THREADS, sm);
    } else {
        cudaFuncSetAttribute(k_enumerate_fast<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_enumerate_fast<false, 3>, FAST_THREADS, sm);
    }
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

void launch_enumerate_fast(const ProblemView& pv, const FastTables& ft, const EnumWork& w, int grid, bool framed,
                           int slots_needed, cudaStream_t s) {
    const int slots = fast_slots(slots_needed, framed);
    const size_t sm = fast_smem_bytes(w.topk, framed, slots);
    if (framed) k_enumerate_fast<true, 3><<<grid, FAST_THREADS, sm, s>>>(pv, ft, w);
    else if (slots == 2) k_enumerate_fast<false, 2><<<grid, FAST_THREADS, sm, s>>>(pv, ft, w);
    else k_enumerate_fast<false, 3><<<grid, FAST_THREADS, sm, s>>>(pv, ft, w);
}

// ---------------------------------------------------------------------- explicit items (fp32)
__global__ void __launch_bounds__(FAST_THREADS) k_score_items_fast(ProblemView pv, FastTables ft,
                                                                   const ScoreItem* __restrict__ items, int64_t n,
                                                                   double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSmem* ws = pin_shared(reinterpret_cast<WarpSmem*>(smem_raw) + (threadIdx.x >> 5));
    int lane = threadIdx.x & 31;
    asm volatile("" : "+r"(lane));
    const int64_t idx = (int64_t)blockIdx.x * FAST_WARPS + (threadIdx.x >> 5);
    if (idx >= n) return;
    warp_init_guards(ws, lane);
    const ScoreItem it = items[idx];
    int ok = 1;
    if (lane == 0) ok = build_program_forest(it.k, it.parent, it.nodes, program_index(ft.tb[it.p * pv.C + it.c]), ft.lut, &ws->prog) ? 1 : 0;
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) { if (lane == 0) out[idx] = CUDART_NAN; return; }
    unsigned s0 = 0, s1 = 0, s2 = 0;
    const double score = warp_eval_tree<false, WarpSmem>(pv, ft, it.p, it.c, ws, lane, &s0, &s1, &s2);
    if (lane == 0) out[idx] = score;
}

void launch_score_items_fast(const ProblemView& pv, const FastTables& ft, const ScoreItem* d_items, int64_t n,
                             double* d_out, cudaStream_t s) {
    if (n <= 0) return;
    cudaFuncSetAttribute(k_score_items_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes(1));
    const unsigned blocks = (unsigned)((n + FAST_WARPS - 1) / FAST_WARPS);
    k_score_items_fast<<<blocks, FAST_THREADS, fast_smem_bytes(1), s>>>(pv, ft, d_items, n, d_out);
}

}  // namespace bamse
