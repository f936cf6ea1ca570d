// Tree index -> parent vector -> evaluation program + canonizeF weight.
// Host and device share this code (bamse_decode_tree uses the host instance).
#pragma once
#include "common.cuh"

namespace bamse {

// index = root + K * code, code = K-2 Pruefer digits base K, least significant first.
// Definition: oracle/bamse_oracle.py decode_tree.  Search space = all K^(K-1) labeled
// rooted trees, i.e. what the reference's branch-and-bound walks (clustering.py:116-148).
__host__ __device__ inline void decode_tree(int K, uint64_t index, int8_t* parent) {
    if (K == 1) { parent[0] = -1; return; }
    const int root = (int)(index % (uint64_t)K);
    uint64_t code = index / (uint64_t)K;
    int8_t seq[KMAX];
    int8_t deg[KMAX];
    uint16_t adj[KMAX];  // adjacency bit masks
    _Pragma("unroll 1") for (int v = 0; v < K; ++v) { deg[v] = 1; adj[v] = 0; }
    _Pragma("unroll 1") for (int i = 0; i < K - 2; ++i) {
        seq[i] = (int8_t)(code % (uint64_t)K);
        code /= (uint64_t)K;
        deg[seq[i]]++;
    }
    _Pragma("unroll 1") for (int i = 0; i < K - 2; ++i) {
        int leaf = 0;
        _Pragma("unroll 1") while (deg[leaf] != 1) ++leaf;  // smallest label with degree 1
        const int s = seq[i];
        adj[leaf] |= (uint16_t)(1u << s);
        adj[s] |= (uint16_t)(1u << leaf);
        deg[leaf]--;
        deg[s]--;
    }
    int u = -1, w = -1;
    _Pragma("unroll 1") for (int v = 0; v < K; ++v)
        if (deg[v] == 1) { if (u < 0) u = v; else w = v; }
    adj[u] |= (uint16_t)(1u << w);
    adj[w] |= (uint16_t)(1u << u);
    // orient away from the root
    _Pragma("unroll 1") for (int v = 0; v < K; ++v) parent[v] = -2;
    parent[root] = -1;
    int8_t stack[KMAX];
    int sp = 0;
    stack[sp++] = (int8_t)root;
    _Pragma("unroll 1") while (sp > 0) {
        const int n = stack[--sp];
        uint16_t nb = adj[n];
        _Pragma("unroll 1") while (nb) {
#ifdef __CUDA_ARCH__
            const int x = __ffs((int)nb) - 1;
#else
            const int x = __builtin_ctz((unsigned)nb);
#endif
            nb &= (uint16_t)(nb - 1);
            if (parent[x] == -2) { parent[x] = (int8_t)n; stack[sp++] = (int8_t)x; }
        }
    }
}

// Builds the evaluation program and the canonizeF weight of `parent[0..k)`.
// nodes[v] = table row of local node v (identity when nodes == nullptr).
// Returns false when the vector is not a rooted tree (exactly one -1, no cycles).
//
// canonizeF (tree_tools.py:47-63): a leaf is '[]' with weight 1; an internal node's
// weight is (product of child weights) x (product of RUN LENGTHS of consecutive equal
// child strings, children in index order, unsorted); its string is '[' + sorted child
// strings + ']'.  Strings are kept as left-aligned bit strings ('[' = 0, ']' = 1); no
// balanced string is a proper prefix of another, so comparing the left-aligned
// integers orders them exactly like Python compares the ASCII strings ('[' < ']').
__host__ __device__ inline bool build_program(int k, const int8_t* parent, const int8_t* nodes,
                                              TreeProgram* prog) {
    int root = -1;
    _Pragma("unroll 1") for (int v = 0; v < k; ++v) {
        if (parent[v] == -1) { if (root >= 0) return false; root = v; }
        else if (parent[v] < 0 || parent[v] >= k || parent[v] == v) return false;
    }
    if (root < 0) return false;

    int8_t nchild[KMAX];
    _Pragma("unroll 1") for (int v = 0; v < k; ++v) nchild[v] = 0;
    _Pragma("unroll 1") for (int v = 0; v < k; ++v) if (parent[v] >= 0) nchild[parent[v]]++;

    // iterative DFS; frame = (node, next child candidate index to scan from)
    int8_t st_node[KMAX], st_next[KMAX], st_seen[KMAX];
    uint32_t canon[KMAX];       // left-aligned bit string of the finished subtree
    int8_t canon_len[KMAX];
    int32_t weight[KMAX];
    int n_ops = 0, n_leaves = 0, visited = 0;
    int sp = 0;
    st_node[0] = (int8_t)root; st_next[0] = 0; st_seen[0] = 0; sp = 1;
    _Pragma("unroll 1") while (sp > 0) {
        const int v = st_node[sp - 1];
        if (nchild[v] == 0) {                         // leaf
            if (n_ops >= MAX_OPS) return false;
            prog->op[n_ops] = OP_LEAF; prog->arg[n_ops] = nodes ? nodes[v] : (int8_t)v; ++n_ops;
            canon[v] = 0x40000000u;                   // "[]" = bits 0,1 left aligned
            canon_len[v] = 2; weight[v] = 1;
            ++n_leaves; ++visited; --sp;
        } else {
            int c = st_next[sp - 1];
            _Pragma("unroll 1") while (c < k && parent[c] != v) ++c;
            if (c < k) {                              // descend into next child
                st_next[sp - 1] = (int8_t)(c + 1);
                if (sp >= KMAX) return false;         // cycle guard
                st_node[sp] = (int8_t)c; st_next[sp] = 0; st_seen[sp] = 0; ++sp;
                continue;
            }
            // all children done: fold canon strings / weights
            uint32_t cs[KMAX]; int8_t cl[KMAX]; int nc = 0;
            int32_t w = 1; int run = 1; uint32_t prev = 0; int8_t prevlen = 0;
            _Pragma("unroll 1") for (int ch = 0; ch < k; ++ch) if (parent[ch] == v) {
                w *= weight[ch];
                if (nc > 0 && canon[ch] == prev && canon_len[ch] == prevlen) ++run;
                else { w *= run; run = 1; }
                prev = canon[ch]; prevlen = canon_len[ch];
                // insertion sort ascending
                int i = nc++;
                _Pragma("unroll 1") while (i > 0 && cs[i - 1] > canon[ch]) { cs[i] = cs[i - 1]; cl[i] = cl[i - 1]; --i; }
                cs[i] = canon[ch]; cl[i] = canon_len[ch];
            }
            w *= run;
            uint32_t s = 0; int len = 1;              // leading '[' = 0 bit
            _Pragma("unroll 1") for (int i = 0; i < nc; ++i) { s |= cs[i] >> len; len += cl[i]; }
            s |= 1u << (31 - len); ++len;             // trailing ']' = 1 bit
            canon[v] = s; canon_len[v] = (int8_t)len; weight[v] = w;
            if (n_ops >= MAX_OPS) return false;
            prog->op[n_ops] = OP_ADDD; prog->arg[n_ops] = nodes ? nodes[v] : (int8_t)v; ++n_ops;
            ++visited; --sp;
        }
        // tell the parent frame that one more child message is on the stack
        if (sp > 0) {
            if (n_ops >= MAX_OPS) return false;
            prog->op[n_ops] = (st_seen[sp - 1] == 0) ? OP_SCAN : OP_CONV; prog->arg[n_ops] = 0; ++n_ops;
            st_seen[sp - 1]++;
        }
    }
    if (visited != k) return false;                   // disconnected (cycle elsewhere)
    prog->n_ops = (int8_t)n_ops; prog->k = (int8_t)k; prog->n_leaves = (int8_t)n_leaves; prog->pad = 0;
    prog->scre = weight[root];
    return true;
}

}  // namespace bamse
