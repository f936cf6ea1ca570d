// Inside x outside factorisation of the fp32 production path ("pairs").
//
// The score of a tree is multilinear in the messages of its subtrees.  Cut a tree at a node v into a PACK
// (some of v's child subtrees: a labeled rooted forest F on the node set S, |S| <= s) and the CONTEXT (the
// rest: a rooted tree tau on R = all \ S with the marked node v).  Then
//
//     score = log sum_z g_F[z] * t_(tau,v)[z]
//
// with g_F the pack's inside message (row of the sub-forest tables FG, forest.cuh) and t_(tau,v) the context's
// OUTSIDE message: with O_v the outside message at v (O_root = 1/L), A_v[y] = sum_{x>=y} O_v[x] d_v[x] and H the
// forest of v's children that stay in the context,
//
//     t_(tau,v)[z] = e^p0 / L^2 * sum_j g_H[j] A_v[z+j]          (H empty: e^p0 / L * A_v[z])
//     O_w          = t_(tau minus subtree(w), v)                   for a child w of v
//
// so a context row follows from the row of a context with FEWER nodes (tau minus subtree(v), marked parent(v))
// by one pointwise product, one suffix sum and one correlation: the same level-by-level build as the inside
// tables.  Per (parameter set, clustering, sample) the tables hold every forest with <= s nodes and every
// context with <= r nodes; a tree whose best cut has K - r <= |S| <= s ("regular") costs ONE windowed dot product
// of two table rows per sample instead of convolutions.  K = 7 (s = 4, r = 4): 7 826 rows in use serve all 117 649
// trees; K = 10 (s = 5, r = 6): 2.7 M rows serve 98 % of the 10^9 trees.  The other trees ("irregular") run
// the table-driven programs of forest.cuh.
//
// Context row index inside the block of one (p, c, m):  base[q] + rank(R) * q^q + tree_index(tau) * q + local(v),
// q = |R|, rank = combinatorial number system over the sorted labels, tree_index = root + q * Pruefer code of tau
// written with local labels (the K^(K-1) indexing of tree_decode.cuh at K = q).
#pragma once
#include "forest.cuh"

namespace bamse {

constexpr int PR_RMAX = 6;                 // largest context the index math supports

__host__ __device__ inline uint32_t ctx_count(int q) {        // q^q: rooted trees on q labeled nodes x marked node
    return q <= 1 ? 1u : q == 2 ? 4u : q == 3 ? 27u : q == 4 ? 256u : q == 5 ? 3125u : 46656u;
}

struct PairGeom {
    int K, s, r;
    uint32_t base[PR_RMAX + 2];            // base[q] = first row of the q-node contexts, base[r+1] = nt
    uint32_t nt;                           // context rows per (p, c, m)
};

__host__ __device__ inline PairGeom make_pair_geom(int K, int s, int r) {
    PairGeom g;
    g.K = K; g.s = s; g.r = r;
    uint32_t acc = 0;
    g.base[0] = 0;
    for (int q = 1; q <= PR_RMAX + 1; ++q) {
        g.base[q] = acc;
        if (q <= r && q <= K) acc += (uint32_t)binom_small(K, q) * ctx_count(q);
    }
    g.nt = acc;
    return g;
}

// inverse of decode_tree for q <= PR_RMAX local labels: parent vector -> root + q * Pruefer code
__host__ __device__ inline uint32_t encode_tree(int q, const int8_t* par) {
    if (q <= 1) return 0;
    uint16_t adj[PR_RMAX];
    int8_t deg[PR_RMAX];
    int root = 0;
    _Pragma("unroll 1") for (int v = 0; v < q; ++v) { adj[v] = 0; deg[v] = 0; }
    _Pragma("unroll 1") for (int v = 0; v < q; ++v) {
        if (par[v] < 0) { root = v; continue; }
        adj[v] |= (uint16_t)(1u << par[v]); adj[par[v]] |= (uint16_t)(1u << v);
        deg[v]++; deg[par[v]]++;
    }
    uint32_t code = 0, mul = 1;
    _Pragma("unroll 1") for (int i = 0; i < q - 2; ++i) {
        int leaf = 0;
        _Pragma("unroll 1") while (deg[leaf] != 1) ++leaf;          // smallest label with degree 1
        const int nb = BAMSE_CTZ(adj[leaf]);
        code += (uint32_t)nb * mul;
        mul *= (uint32_t)q;
        adj[nb] &= (uint16_t)~(1u << leaf);
        adj[leaf] = 0; deg[leaf] = 0; deg[nb]--;
    }
    return (uint32_t)root + (uint32_t)q * code;
}

// Row of the context on the node set `mask` (closed under gp: every node's parent is in the set, except the
// root's) with the marked node v.
__host__ __device__ inline uint32_t ctx_row(const PairGeom& g, uint32_t mask, const int8_t* gp, int v) {
    const int q = BAMSE_POPC(mask);
    int8_t par[PR_RMAX];
    int rank = 0, i = 0;
    for (uint32_t mm = mask; mm; mm &= mm - 1, ++i) {
        const int l = BAMSE_CTZ(mm);
        rank += binom_small(l, i + 1);
        const int p = gp[l];
        par[i] = (p >= 0 && ((mask >> p) & 1u)) ? (int8_t)BAMSE_POPC(mask & ((1u << p) - 1u)) : (int8_t)-1;
    }
    return g.base[q] + (uint32_t)rank * ctx_count(q) + encode_tree(q, par) * (uint32_t)q +
           (uint32_t)BAMSE_POPC(mask & ((1u << v) - 1u));
}

// The cut of a full tree (parent vector over the labels 0 .. K-1): over all nodes v and all subsets of v's child
// subtrees, the largest pack with <= s nodes (exact subset sums; the first v in label order wins ties, children
// are dropped from the highest label down).  Returns false ("irregular") when even that pack leaves a context of
// more than r nodes.
__host__ __device__ inline bool classify_tree(int K, const int8_t* parent, const ForestIndex& fi,
                                              const int16_t* __restrict__ lut, const PairGeom& g, uint32_t* fg_row,
                                              uint32_t* t_row, uint32_t* pack_mask = nullptr) {
    if (K < 2) return false;
    int8_t size[KMAX], depth[KMAX], order[KMAX];
    uint16_t sub[KMAX];
    _Pragma("unroll 1") for (int l = 0; l < K; ++l) {
        int d = 0, w = l;
        _Pragma("unroll 1") while (parent[w] >= 0) { w = parent[w]; ++d; }
        depth[l] = (int8_t)d; size[l] = 1; sub[l] = (uint16_t)(1u << l);
        int j = l;                                     // insertion sort by decreasing depth
        _Pragma("unroll 1") while (j > 0 && depth[order[j - 1]] < d) { order[j] = order[j - 1]; --j; }
        order[j] = (int8_t)l;
    }
    _Pragma("unroll 1") for (int i = 0; i < K; ++i) {
        const int l = order[i];
        if (parent[l] >= 0) { size[parent[l]] += size[l]; sub[parent[l]] |= sub[l]; }
    }
    const int s = g.s;
    const uint32_t keep = (2u << s) - 1u;              // sums 0 .. s
    int best = 0, best_v = -1;
    _Pragma("unroll 1") for (int v = 0; v < K; ++v) {
        uint32_t reach = 1u;
        _Pragma("unroll 1") for (int c = 0; c < K; ++c)
            if (parent[c] == v) reach |= (reach << size[c]) & keep;
        const int top = 31 - BAMSE_CLZ(reach);
        if (top > best) { best = top; best_v = v; }
    }
    if (best_v < 0 || K - best > g.r) return false;
    // which children: walk the children from the highest label down, dropping a child whenever the remaining
    // target is still reachable without it
    uint32_t pack = 0;
    {
        int8_t kids[KMAX];
        uint32_t before[KMAX];                         // sums reachable with the children in front of kids[i]
        int nk = 0;
        uint32_t reach = 1u;
        _Pragma("unroll 1") for (int c = 0; c < K; ++c)
            if (parent[c] == best_v) { kids[nk] = (int8_t)c; before[nk] = reach; ++nk; reach |= (reach << size[c]) & keep; }
        int target = best;
        _Pragma("unroll 1") for (int i = nk - 1; i >= 0 && target > 0; --i) {
            if ((before[i] >> target) & 1u) continue;  // reachable without this child
            pack |= sub[kids[i]];
            target -= size[kids[i]];
        }
    }
    if (pack_mask) *pack_mask = pack;
    *fg_row = (uint32_t)forest_row(fi, lut, pack, parent);
    *t_row = ctx_row(g, ((1u << K) - 1u) & ~pack, parent, best_v);
    return true;
}

// Node-set sharding of a (parameter set, clustering) that is split over all ranks: the rank that scores the
// regular trees whose pack sits on the node set `mask` -- and therefore builds the top-level inside rows on that
// set and the top-level context rows on its complement (top-level rows are never ingredients of other rows; the
// lower levels are built everywhere).  Round robin over the combinatorial rank of the set among the sets of its size.
__host__ __device__ inline int node_set_owner(uint32_t mask, int world) {
    int rank = 0, i = 0;
    for (uint32_t mm = mask; mm; mm &= mm - 1, ++i) rank += binom_small(BAMSE_CTZ(mm), i + 1);
    return rank % world;
}

// How a context row is produced (host-generated, one array per (K, s, r)).
struct CtxRecipe {
    int32_t parent_row;    // context of the marked node's parent (fewer nodes), -1: the marked node is the root
    int32_t h_row;         // FG row of the marked node's children forest inside the context, -1: none
    int8_t v;              // label of the marked node
    int8_t pad;
    uint16_t mask;         // node set of the context
};
static_assert(sizeof(CtxRecipe) == 12, "recipe layout");

// One entry of a per-K pair table (indexed by tree index): inside row (20 bits) | compact context row (22 bits) << 20
// | scre (22 bits) << 42; 0 marks an irregular tree (scre >= 1 otherwise).  K = 10: 354 907 inside rows, 2.4 M
// context rows in use, scre <= 9! = 362 880.
constexpr uint32_t PAIR_FG_BITS = 20, PAIR_T_BITS = 22, PAIR_SCRE_BITS = 22;
__host__ __device__ inline uint64_t pair_entry(uint32_t fg_row, uint32_t t_row, uint32_t scre) {
    return (uint64_t)fg_row | ((uint64_t)t_row << PAIR_FG_BITS) | ((uint64_t)scre << (PAIR_FG_BITS + PAIR_T_BITS));
}
__host__ __device__ inline void pair_unpack(uint64_t e, uint32_t* fg_row, uint32_t* t_row, uint32_t* scre) {
    *fg_row = (uint32_t)e & ((1u << PAIR_FG_BITS) - 1u);
    *t_row = (uint32_t)(e >> PAIR_FG_BITS) & ((1u << PAIR_T_BITS) - 1u);
    *scre = (uint32_t)(e >> (PAIR_FG_BITS + PAIR_T_BITS));
}

}  // namespace bamse
