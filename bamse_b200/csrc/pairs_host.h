// Host-side generator of the context recipes (pairs.cuh): how every context row follows from a context with
// fewer nodes and an inside-table row.  Data independent, one array per (K, s, r).
#pragma once
#include <vector>

#include "forest_host.h"
#include "pairs.cuh"

namespace bamse {

// recipes of all g.nt rows, in row order; fi must be the inside geometry the rows' h_row refer to (fi.s >= g.r - 1)
inline std::vector<CtxRecipe> make_ctx_recipes(const ForestIndex& fi, const PairGeom& g) {
    const ForestLut& fl = forest_lut();
    std::vector<CtxRecipe> out((size_t)g.nt);
    const int K = g.K;
    for (uint32_t mask = 1; mask < (1u << K); ++mask) {
        const int q = BAMSE_POPC(mask);
        if (q > g.r) continue;
        int lab[PR_RMAX], i = 0;
        for (uint32_t mm = mask; mm; mm &= mm - 1) lab[i++] = BAMSE_CTZ(mm);
        const uint64_t n_trees = ipow_u64(q, q - 1);
        for (uint64_t idx = 0; idx < n_trees; ++idx) {
            int8_t par[KMAX], gp[KMAX];
            decode_tree(q, idx, par);
            for (int v = 0; v < KMAX; ++v) gp[v] = -1;
            for (int j = 0; j < q; ++j) gp[lab[j]] = par[j] >= 0 ? (int8_t)lab[par[j]] : (int8_t)-1;
            for (int j = 0; j < q; ++j) {
                const int v = lab[j];
                uint32_t desc = 0;                     // proper descendants of v inside the context
                for (int t = 0; t < q; ++t) {
                    int w = lab[t];
                    if (w == v) continue;
                    while (gp[w] >= 0 && gp[w] != v) w = gp[w];
                    if (gp[w] == v) desc |= 1u << lab[t];
                }
                CtxRecipe r;
                r.v = (int8_t)v; r.pad[0] = r.pad[1] = r.pad[2] = 0;
                r.h_row = desc ? forest_row(fi, fl.lut.data(), desc, gp) : -1;
                r.parent_row = gp[v] >= 0 ? (int32_t)ctx_row(g, mask & ~(desc | (1u << v)), gp, gp[v]) : -1;
                out[(size_t)ctx_row(g, mask, gp, v)] = r;
            }
        }
    }
    return out;
}

}  // namespace bamse
