// Host-side generator of the context recipes (pairs.cuh): how every context row follows from a context with
// fewer nodes and an inside-table row.  Data independent, one array per (K, s, r).
#pragma once
#include <vector>

#include "forest_host.h"
#include "pairs.cuh"

namespace bamse {

// recipe of ONE context row (full index space): unrank the node set, decode the tree, find the marked node's
// children forest and the context its outside message comes from
inline CtxRecipe ctx_recipe_of_row(const ForestIndex& fi, const PairGeom& g, uint32_t row) {
    const ForestLut& fl = forest_lut();
    int q = 1;
    while (q < g.r && row >= g.base[q + 1]) ++q;
    const uint32_t local = row - g.base[q];
    uint32_t rank = local / ctx_count(q);
    const uint32_t rest = local % ctx_count(q);
    const uint64_t idx = rest / (uint32_t)q;
    const int j = (int)(rest % (uint32_t)q);
    int lab[PR_RMAX];
    uint32_t mask = 0;
    for (int i = q - 1; i >= 0; --i) {               // combinatorial number system: rank = sum binom(lab[i], i + 1)
        int l = i;
        while (l + 1 < g.K && (uint32_t)binom_small(l + 1, i + 1) <= rank) ++l;
        lab[i] = l;
        rank -= (uint32_t)binom_small(l, i + 1);
        mask |= 1u << l;
    }
    int8_t par[KMAX], gp[KMAX];
    decode_tree(q, idx, par);
    for (int v = 0; v < KMAX; ++v) gp[v] = -1;
    for (int t = 0; t < q; ++t) gp[lab[t]] = par[t] >= 0 ? (int8_t)lab[par[t]] : (int8_t)-1;
    const int v = lab[j];
    uint32_t desc = 0;                                 // proper descendants of v inside the context
    for (int t = 0; t < q; ++t) {
        int w = lab[t];
        if (w == v) continue;
        while (gp[w] >= 0 && gp[w] != v) w = gp[w];
        if (gp[w] == v) desc |= 1u << lab[t];
    }
    CtxRecipe r;
    r.v = (int8_t)v; r.pad = 0; r.mask = (uint16_t)mask;
    r.h_row = desc ? forest_row(fi, fl.lut.data(), desc, gp) : -1;
    r.parent_row = gp[v] >= 0 ? (int32_t)ctx_row(g, mask & ~(desc | (1u << v)), gp, gp[v]) : -1;
    return r;
}

// recipes of all g.nt rows, in row order; fi must be the inside geometry the rows' h_row refer to (fi.s >= g.r - 1)
inline std::vector<CtxRecipe> make_ctx_recipes(const ForestIndex& fi, const PairGeom& g) {
    std::vector<CtxRecipe> out((size_t)g.nt);
    for (uint32_t row = 0; row < g.nt; ++row) out[row] = ctx_recipe_of_row(fi, g, row);
    return out;
}

}  // namespace bamse
