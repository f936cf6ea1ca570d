// Sub-forest memoisation of the fp32 production path (design hint: the memo key "sorted node
// set + shape" of the reference's dead prufer_treeset, clustering.py:299-318).
//
// Every message of the sum-rule integration is a function of a LABELED ROOTED FOREST on a subset of the
// clusters: G(F) = (*)_{tau in F} E(tau), E(tau) = D[root] + SG(children forest), SG(F) = scan(G(F)) + p0 - log L.
// The K^(K-1) trees of a clustering share their small sub-forests massively, so per (parameter set,
// clustering, sample) the messages of ALL forests with <= s nodes are built once, level by level, with the
// production convolution ("message tables" FG / FSG), and for every root r and forest F with <= sh nodes the
// finishing vector FH[r][F][i] = 1/L^2 sum_j d_r[i+j] sg_F[j] (the dual of "convolve with SG(F), add D[r],
// integrate").  A tree's evaluation program then only convolves what is specific to it: the children of a
// node whose subtrees have <= s nodes are packed into table forests (one table load each), one of the root's
// packs disappears into the finishing table, and only nodes with larger subtrees cost convolutions.
// Convolutions per tree: K = 6: 0.65 -> 0 (s = 4), K = 7: 0.98 -> 0.24, K = 8: 1.30 -> 0.46, K = 10: 1.93 -> 1.03.
//
// Row index of a forest inside the table block of one (p, c, m):  base[n] + rank(node set) * N_n + shape,
// n = number of nodes, rank = combinatorial number system over the sorted labels, N_n = (n+1)^(n-1) labeled
// rooted forests on n nodes, shape = dense index of the parent map written with local indices (a LUT over the
// n-digit base-(n+1) code).  Rows 0 .. K-1 are the single nodes: FG = leaf message EL, FSG = scanned leaf SEL.
#pragma once
#include "common.cuh"
#include "tree_decode.cuh"

namespace bamse {

constexpr int FT_SMAX = 5;                 // largest forest the index math supports
constexpr int FT_LUT_SIZE = 8476;          // sum_{n=1..5} (n+1)^n codes

#ifdef __CUDA_ARCH__
#define BAMSE_POPC(x) __popc((unsigned)(x))
#define BAMSE_CTZ(x) (__ffs((int)(x)) - 1)
#define BAMSE_CLZ(x) __clz((int)(x))
#else
#define BAMSE_POPC(x) __builtin_popcount((unsigned)(x))
#define BAMSE_CTZ(x) __builtin_ctz((unsigned)(x))
#define BAMSE_CLZ(x) __builtin_clz((unsigned)(x))
#endif

__host__ __device__ inline int forest_count(int n) {          // (n+1)^(n-1)
    return n <= 1 ? 1 : n == 2 ? 3 : n == 3 ? 16 : n == 4 ? 125 : 1296;
}
__host__ __device__ inline int forest_lut_offset(int n) {     // sum_{j<n} (j+1)^j
    return n <= 1 ? 0 : n == 2 ? 2 : n == 3 ? 11 : n == 4 ? 75 : 700;
}
__host__ __device__ inline int binom_small(int n, int k) {
    if (k < 0 || k > n) return 0;
    int r = 1;
    for (int i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
    return r;
}

// Table geometry of one clustering: K clusters, message tables for forests of <= s nodes, finishing tables
// for forests of <= sh nodes (sh <= s).
struct ForestIndex {
    int K, s, sh;
    int base[FT_SMAX + 2];   // base[n] = first row of the n-node forests, base[s+1] = nf
    int nf;                  // rows of FG / FSG per (p, c, m)
    int nfh;                 // rows of FH per (p, c, m, root) = base[sh+1]
};

__host__ __device__ inline ForestIndex make_forest_index(int K, int s, int sh) {
    ForestIndex fi;
    fi.K = K; fi.s = s; fi.sh = sh;
    int acc = 0;
    fi.base[0] = 0;
    for (int n = 1; n <= FT_SMAX + 1; ++n) {
        fi.base[n] = acc;
        if (n <= s) acc += binom_small(K, n) * forest_count(n);
    }
    fi.nf = fi.base[s + 1];
    fi.nfh = fi.base[sh + 1];
    return fi;
}

// Row of the forest on the node set `mask` whose parent map is gp[label] (a parent outside the mask, or -1,
// makes the node a root of the forest).
__host__ __device__ inline int forest_row(const ForestIndex& fi, const int16_t* __restrict__ lut, uint32_t mask,
                                          const int8_t* gp) {
    const int n = BAMSE_POPC(mask);
    int rank = 0, code = 0, mul = 1, i = 0;
    for (uint32_t mm = mask; mm; mm &= mm - 1, ++i) {
        const int l = BAMSE_CTZ(mm);
        rank += binom_small(l, i + 1);
        const int p = gp[l];
        const int d = (p >= 0 && ((mask >> p) & 1u)) ? BAMSE_POPC(mask & ((1u << p) - 1u)) + 1 : 0;
        code += d * mul;
        mul *= n + 1;
    }
    return fi.base[n] + rank * forest_count(n) + lut[forest_lut_offset(n) + code];
}

// ---------------------------------------------------------------------------------- fast programs
enum : int8_t {
    FOP_PUSH_G = 0,     // push FG[arg]    message of a table forest
    FOP_PUSH_SG = 1,    // push FSG[arg]   scanned message (the forest carries its parent's prefix sum)
    FOP_SCAN = 2,       // top = log2cumsum(top) + p0 - log L
    FOP_CONV = 3,       // b = pop; top = top (*) b   (in place)
    FOP_ADDD = 4,       // top += D2[arg]
};
enum : int8_t {
    FIN_LSE = 0,        // log2sumexp(top) - log2 L                 (trap_int(...)[-1], clustering.py:385)
    FIN_DOT_D = 1,      // log2 sum 2^(top + D2[r]) - log2 L        all of the root's children in one table forest
    FIN_DOT_SD = 2,     // log2 sum 2^(top + SD2[r])                no finishing forest at the root
    FIN_DOT_FH = 3,     // log2 sum 2^(top + FH[r][arg])            finishing forest `arg`, top = the other children
};

constexpr int FAST_MAX_OPS = 20;           // brute-forced: <= 19 for K <= 12 with s >= 2
struct alignas(16) FastProgram {
    uint8_t n_ops, k, n_leaves, fin;
    int32_t scre;                          // canonizeF 'scre'                    (tree_tools.py:47-63)
    uint16_t fin_arg;                      // FIN_DOT_FH: row of the finishing forest
    uint8_t fin_r;                         // root label
    uint8_t max_sp;                        // live vectors the program needs
    uint8_t op4[FAST_MAX_OPS / 2];         // two 4-bit opcodes per byte
    uint16_t arg[FAST_MAX_OPS];
    uint8_t pad[2];
};
static_assert(sizeof(FastProgram) == 64, "program table rows are fetched as 16 words");

__host__ __device__ inline int fast_op(const FastProgram& p, int i) { return (p.op4[i >> 1] >> ((i & 1) * 4)) & 15; }

// Builds the table-driven program of a tree (k nodes, parent vector in local indices, nodes[v] = cluster id of
// local node v or identity) for the table geometry `fi`.  Same validity rules as build_program.
__host__ __device__ inline bool build_program_forest(int k, const int8_t* parent, const int8_t* nodes,
                                                     const ForestIndex& fi, const int16_t* __restrict__ lut,
                                                     FastProgram* prog) {
    TreeProgram ref;                                   // canon weight / validation in the reference order
    if (!build_program(k, parent, nodes, &ref)) return false;
    // label space: everything below is indexed by cluster id
    int8_t gp[KMAX], size[KMAX], depth[KMAX], order[KMAX];
    uint16_t sub[KMAX];
    uint32_t present = 0;
    int root = 0;
    _Pragma("unroll 1") for (int v = 0; v < KMAX; ++v) { gp[v] = -2; size[v] = 0; sub[v] = 0; }
    _Pragma("unroll 1") for (int v = 0; v < k; ++v) {
        const int l = nodes ? nodes[v] : v;
        if (l < 0 || l >= KMAX || ((present >> l) & 1u)) return false;
        present |= 1u << l;
        gp[l] = parent[v] >= 0 ? (int8_t)(nodes ? nodes[parent[v]] : parent[v]) : (int8_t)-1;
        if (parent[v] < 0) root = l;
    }
    int n_lab = 0;
    _Pragma("unroll 1") for (uint32_t mm = present; mm; mm &= mm - 1) {
        const int l = BAMSE_CTZ(mm);
        int d = 0, w = l;
        _Pragma("unroll 1") while (gp[w] >= 0) { w = gp[w]; ++d; }
        depth[l] = (int8_t)d; size[l] = 1; sub[l] = (uint16_t)(1u << l);
        int j = n_lab++;                               // insertion sort by decreasing depth
        _Pragma("unroll 1") while (j > 0 && depth[order[j - 1]] < d) { order[j] = order[j - 1]; --j; }
        order[j] = (int8_t)l;
    }
    _Pragma("unroll 1") for (int i = 0; i < n_lab; ++i) {
        const int l = order[i];
        if (gp[l] >= 0) { size[gp[l]] += size[l]; sub[gp[l]] |= sub[l]; }
    }
    const int s = fi.s, sh = fi.sh;
    int n_ops = 0, sp = 0, max_sp = 0;
    uint8_t fin = FIN_LSE, fin_r = (uint8_t)root;
    uint16_t fin_arg = 0;
    _Pragma("unroll 1") for (int i = 0; i < FAST_MAX_OPS / 2; ++i) prog->op4[i] = 0;
    bool ok = true;
    auto emit = [&](int op, int arg) {
        if (n_ops >= FAST_MAX_OPS) { ok = false; return; }
        prog->op4[n_ops >> 1] |= (uint8_t)(op << ((n_ops & 1) * 4));
        prog->arg[n_ops] = (uint16_t)arg;
        ++n_ops;
        if (op == FOP_PUSH_G || op == FOP_PUSH_SG) { ++sp; if (sp > max_sp) max_sp = sp; }
        else if (op == FOP_CONV) --sp;
    };
    if (n_lab == 1) {
        emit(FOP_PUSH_G, root);                        // a single node: its leaf message, plain log-sum-exp
    } else {
        int8_t st_node[KMAX], st_stage[KMAX], st_nf[KMAX];
        uint16_t st_used[KMAX];
        int fsp = 1;
        st_node[0] = (int8_t)root; st_stage[0] = 0; st_nf[0] = 0; st_used[0] = 0;
        _Pragma("unroll 1") while (fsp > 0 && ok) {
            const int f = fsp - 1, v = st_node[f];
            if (st_stage[f] == 1) {                    // a big child's message is on the stack
                if (st_nf[f] > 0) emit(FOP_CONV, 0);
                st_nf[f]++;
                st_stage[f] = 0;
                continue;
            }
            // next unevaluated child with more than s nodes: largest first (keeps the stack shallow)
            int best = -1;
            _Pragma("unroll 1") for (uint32_t mm = present; mm; mm &= mm - 1) {
                const int c = BAMSE_CTZ(mm);
                if (gp[c] == v && size[c] > s && !((st_used[f] >> c) & 1) && (best < 0 || size[c] > size[best])) best = c;
            }
            if (best >= 0) {
                st_used[f] |= (uint16_t)(1u << best);
                st_stage[f] = 1;
                st_node[fsp] = (int8_t)best; st_stage[fsp] = 0; st_nf[fsp] = 0; st_used[fsp] = 0; ++fsp;
                continue;
            }
            const bool is_root = f == 0;
            uint32_t taken = 0, hmask = 0;
            if (is_root) {                             // the finishing forest: greedy fill up to sh nodes
                int hsize = 0;
                _Pragma("unroll 1") for (int sz = sh; sz >= 1; --sz)
                    _Pragma("unroll 1") for (uint32_t mm = present; mm; mm &= mm - 1) {
                        const int c = BAMSE_CTZ(mm);
                        if (gp[c] == v && size[c] == sz && hsize + sz <= sh) { hmask |= sub[c]; hsize += sz; taken |= 1u << c; }
                    }
            }
            // the other small children: first-fit decreasing into table forests of <= s nodes
            uint16_t bmask[KMAX];
            int8_t bsize[KMAX];
            int nb = 0;
            _Pragma("unroll 1") for (int sz = s; sz >= 1; --sz)
                _Pragma("unroll 1") for (uint32_t mm = present; mm; mm &= mm - 1) {
                    const int c = BAMSE_CTZ(mm);
                    if (gp[c] != v || size[c] != sz || ((taken >> c) & 1u)) continue;
                    int b = 0;
                    _Pragma("unroll 1") while (b < nb && bsize[b] + sz > s) ++b;
                    if (b == nb) { bmask[nb] = 0; bsize[nb] = 0; ++nb; }
                    bmask[b] |= sub[c]; bsize[b] += (int8_t)sz;
                }
            _Pragma("unroll 1") for (int b = 0; b < nb; ++b) {
                emit((!is_root && b == 0) ? FOP_PUSH_SG : FOP_PUSH_G, forest_row(fi, lut, bmask[b], gp));
                if (st_nf[f] > 0) emit(FOP_CONV, 0);
                st_nf[f]++;
            }
            if (is_root) {
                if (hmask) {
                    const int row = forest_row(fi, lut, hmask, gp);
                    if (st_nf[f] == 0) { emit(FOP_PUSH_SG, row); fin = FIN_DOT_D; }
                    else { fin = FIN_DOT_FH; fin_arg = (uint16_t)row; }
                } else fin = FIN_DOT_SD;
            } else {
                if (nb == 0) emit(FOP_SCAN, 0);        // no table forest: the accumulated product carries the scan
                emit(FOP_ADDD, v);
            }
            --fsp;
        }
    }
    if (!ok) return false;
    prog->n_ops = (uint8_t)n_ops; prog->k = (uint8_t)k; prog->n_leaves = (uint8_t)ref.n_leaves; prog->fin = fin;
    prog->scre = ref.scre; prog->fin_arg = fin_arg; prog->fin_r = fin_r; prog->max_sp = (uint8_t)max_sp;
    prog->pad[0] = prog->pad[1] = 0;
    return true;
}

// How a table row is produced from rows with fewer nodes (host-generated, one array per (K, s)).
struct ForestRecipe {
    uint16_t a, b;      // kind 1: b = row of the children forest; kind 2: a (*) b
    uint16_t mask;      // node set
    int8_t kind;        // 0 single node (built from D in fp64), 1 single tree: D2[u] + FSG[b], 2 several trees: FG[a] (*) FG[b]
    int8_t u;           // kind 0 / 1: the root's label
};
static_assert(sizeof(ForestRecipe) == 8, "recipe layout");

// Live vector slots a program can need for (K, s) (brute-forced over all tree shapes / sampled above K = 8;
// re-checked by tests/test_host_cpu.py through bamse_debug_program): the in-place convolution keeps <= 2
// vectors alive once every subtree of <= 3 nodes is a table.
__host__ __device__ inline int forest_slots_needed(int K, int s) {
    if (K <= 8) return 2;
    if (s >= 4) return 2;
    if (s >= 3 && K <= 10) return 2;
    return 3;
}

}  // namespace bamse
