// Host-side generators for the sub-forest tables (forest.cuh): the shape LUT and the per-(K, s) recipes that
// tell the table-building kernels how a row follows from rows with fewer nodes.  Data independent.
#pragma once
#include <vector>

#include "forest.cuh"

namespace bamse {

struct ForestLut {
    std::vector<int16_t> lut;                          // code -> shape index, -1 = not a forest
    std::vector<int> codes[FT_SMAX + 1];               // shape index -> code, per n
    ForestLut() : lut(FT_LUT_SIZE, -1) {
        for (int n = 1; n <= FT_SMAX; ++n) {
            int n_codes = 1;
            for (int i = 0; i < n; ++i) n_codes *= n + 1;
            for (int code = 0; code < n_codes; ++code) {
                int d[FT_SMAX], c = code;
                for (int i = 0; i < n; ++i) { d[i] = c % (n + 1); c /= n + 1; }
                bool ok = true;
                for (int i = 0; i < n && ok; ++i) {    // every node must reach a root without a cycle
                    int w = i, steps = 0;
                    while (d[w] != 0 && steps <= n) { w = d[w] - 1; ++steps; }
                    if (steps > n) ok = false;
                }
                if (ok) {
                    lut[forest_lut_offset(n) + code] = (int16_t)codes[n].size();
                    codes[n].push_back(code);
                }
            }
        }
    }
};

inline const ForestLut& forest_lut() {
    static const ForestLut l;
    return l;
}

// recipes of all fi.nf rows, in row order
inline std::vector<ForestRecipe> make_forest_recipes(const ForestIndex& fi) {
    const ForestLut& fl = forest_lut();
    std::vector<ForestRecipe> out((size_t)fi.nf);
    const int K = fi.K;
    for (uint32_t mask = 1; mask < (1u << K); ++mask) {
        const int n = BAMSE_POPC(mask);
        if (n > fi.s) continue;
        int lab[FT_SMAX], i = 0;
        for (uint32_t mm = mask; mm; mm &= mm - 1) lab[i++] = BAMSE_CTZ(mm);
        for (int shape = 0; shape < forest_count(n); ++shape) {
            int code = fl.codes[n][shape];
            int8_t gp[KMAX];
            for (int v = 0; v < KMAX; ++v) gp[v] = -1;
            int n_roots = 0, root0 = -1;
            for (int j = 0; j < n; ++j) {
                const int d = code % (n + 1);
                code /= n + 1;
                gp[lab[j]] = d ? (int8_t)lab[d - 1] : (int8_t)-1;
                if (!d) { ++n_roots; if (root0 < 0) root0 = lab[j]; }
            }
            const int row = forest_row(fi, fl.lut.data(), mask, gp);
            ForestRecipe r;
            r.a = r.b = 0; r.mask = (uint16_t)mask; r.u = -1;
            if (n == 1) { r.kind = 0; r.u = (int8_t)lab[0]; }
            else if (n_roots == 1) {
                r.kind = 1; r.u = (int8_t)root0;
                r.b = (uint16_t)forest_row(fi, fl.lut.data(), mask & ~(1u << root0), gp);
            } else {
                // the component of the smallest label against the rest
                uint32_t comp = 0;
                int top = lab[0];
                while (gp[top] >= 0) top = gp[top];
                for (int j = 0; j < n; ++j) {
                    int w = lab[j];
                    while (gp[w] >= 0) w = gp[w];
                    if (w == top) comp |= 1u << lab[j];
                }
                r.kind = 2;
                r.a = (uint16_t)forest_row(fi, fl.lut.data(), comp, gp);
                r.b = (uint16_t)forest_row(fi, fl.lut.data(), mask & ~comp, gp);
            }
            out[(size_t)row] = r;
        }
    }
    return out;
}

// processing order of the rows of every level: convolution rows (kind 2) first, so that the warps of a
// table-building CTA get rows of the same cost
inline std::vector<uint32_t> make_level_order(const ForestIndex& fi, const std::vector<ForestRecipe>& rec) {
    std::vector<uint32_t> out((size_t)fi.nf);
    for (int n = 1; n <= fi.s; ++n) {
        size_t w = (size_t)fi.base[n];
        for (int pass = 0; pass < 2; ++pass)
            for (int row = fi.base[n]; row < fi.base[n + 1]; ++row)
                if ((rec[(size_t)row].kind == 2) == (pass == 0)) out[w++] = (uint32_t)row;
    }
    return out;
}

// Default table geometry per cluster count: inside tables for forests of <= s nodes, finishing tables of the
// table-driven programs for forests of <= sh nodes, context tables (pairs.cuh) for contexts of <= r nodes
// (r = 0: no pair path).  Chosen by counting the rows in use against the trees they serve and the irregular trees
// left over, for every candidate (s, r), by cutting ALL trees on the GPU (profiles/r02_pair_geometry.txt):
//   K = 6 (3,4): 371 + 846 rows, K = 7 (4,4): 5 005 + 2 821, K = 8 (4,5): 9 738 + 40 552, K = 9 (5,6): 180 507 +
//   262 989 -- all trees regular; K = 10 (5,6): 354 907 + 2 371 810 rows, 2.0 % of the 10^9 trees irregular.
// K >= 11 would need s + r >= 11 (tens of GB per sample): programs only.
inline void default_forest_geometry(int K, int* s, int* sh, int* r = nullptr) {
    int rr = 0;
    if (K <= 1) { *s = 1; *sh = 1; }
    else if (K == 2) { *s = 1; *sh = 1; rr = 1; }
    else if (K == 3) { *s = 2; *sh = 1; rr = 1; }
    else if (K == 4) { *s = 2; *sh = 2; rr = 2; }
    else if (K == 5) { *s = 3; *sh = 2; rr = 3; }
    else if (K == 6) { *s = 3; *sh = 2; rr = 4; }
    else if (K == 7) { *s = 4; *sh = 3; rr = 4; }
    else if (K == 8) { *s = 4; *sh = 3; rr = 5; }
    else if (K <= 10) { *s = 5; *sh = 3; rr = 6; }
    else { *s = 3; *sh = 2; }                         // K = 11, 12
    if (r) *r = rr;
}

// The table-driven programs address rows with 16 bits: they see the inside tables up to s_prog <= s nodes (the
// rows of the smaller forests are a prefix of the table, so both geometries share one table).
inline int program_forest_s(int K, int s) {
    while (s > 1 && make_forest_index(K, s, 1).nf > 65535) --s;
    return s;
}

inline bool reduce_forest_geometry(int* s, int* sh) {
    if (*sh > 1 && *sh >= *s - 1) { --*sh; return true; }
    if (*s > 2) { --*s; if (*sh > *s) *sh = *s; return true; }
    if (*sh > 1) { --*sh; return true; }
    return false;
}

}  // namespace bamse
