// fp32 production scorer: one WARP per tree, vectors kept in shared memory as fp32 log2
// values.  Every message in this computation is a log-concave sequence (binomial
// likelihoods are log-concave; prefix sums, truncated convolutions and pointwise
// products preserve log-concavity), which the convolution exploits:
//
//   * merge path: for log-concave a, b the maximising split (i*, j*) of every output
//     x = i + j walks the merge of the two decreasing slope sequences, so the exact
//     per-output max-shift M_x costs O(L) per convolution instead of an O(L^2) max pass;
//   * windowing: the terms of one output are concave in j, so the ones within THETA bits
//     of M_x form a contiguous window around j*; only those are exponentiated (MUFU.EX2),
//     32 consecutive outputs at a time over the union of their windows (one such pass per
//     half-warp, two adjacent outputs per lane: conv_pass_pair).
//
// Results equal the reference's my_convolve (clustering.py:365-370) up to fp32 rounding:
// dropped terms are below 2^-THETA of the output they belong to.
#pragma once
#include "common.cuh"
#include "tree_decode.cuh"
#include "forest.cuh"
#include "pairs.cuh"

namespace bamse {

constexpr int FAST_WARPS = 4;                       // warps per CTA
constexpr int FAST_THREADS = FAST_WARPS * 32;
constexpr int GUARD = 36;                           // -inf floats in front of every vector (a[x-j], j <= xmax+3)
constexpr int VSTRIDE = GUARD + L + 4;              // floats per vector slot (4 trailing -inf: b[j], j <= 514)
constexpr int FAST_SLOTS = 3;                       // max live vectors with the in-place convolution (brute-forced, K <= 12)
constexpr float THETA = 26.0f;                      // window threshold in bits: terms below 2^-26 of an output's
                                                    // maximum are dropped.  They decay at least geometrically
                                                    // (concavity) with the ratio of the last step inside the
                                                    // window, <= 2^(-26/512): at most 4e-7 relative in the worst
                                                    // case, ~3e-8 on peaked data -- below the rounding of the fp32
                                                    // log values themselves
constexpr float LOG2_L = 9.0f;                      // log2(512)
// framed (linear-domain) path of the convolution
constexpr int FR_B = 64;                            // outputs per framed block (2 per lane)
constexpr int FR_SEG = 128;                         // j's per framed segment
constexpr int FR_A = FR_SEG + FR_B;                 // a-window floats of one segment
constexpr float FR_DEV = 50.0f;                     // allowed deviation of M_x from linear inside a block (bits)
constexpr float FR_MAXR = 30.0f;                    // a factor above 2^FR_MAXR of the centre aborts the frame

// Per (parameter set, clustering): where its tables start and how they are indexed (forest.cuh, pairs.cuh).
struct TableBase {
    uint32_t fg_row;     // first row of FG of (p, c, m = 0); sample m starts fi.nf * m rows later
    uint32_t fsg_row;    // first row of FSG of (p, c, m = 0); sample m starts nfsg rows later
    int32_t nfsg;        // FSG rows per sample: the forests whose scanned message anything reads (a prefix of the rows)
    uint32_t fh_row;     // first row of FH of (p, c, m = 0, root 0); (m, r) starts (m * K + r) * fi.nfh rows later
    uint32_t t_row;      // first row of T of (p, c, m = 0); sample m starts PairK::nt rows later
    int32_t s_prog;      // the table-driven programs pack forests of <= s_prog <= fi.s nodes (16-bit row ids)
    int32_t mine;        // 0: another rank owns this (p, c) under the sharding plan -- no tables here
    int32_t split;       // 1: this (p, c) is split over all ranks (node-set sharding of the pair path, pairs.cuh)
    ForestIndex fi;      // geometry of the inside tables
};

// Per cluster count: the context tables of the pair path (pairs.cuh).  Only the contexts some tree's cut uses
// (and the contexts their recipes refer to) are built: `map` sends a context row of the full index space to its
// compact row, rows are ordered by node count so that a level is a contiguous range.
struct PairK {
    PairGeom g;                    // full index space; g.r == 0: no pair path for this cluster count
    uint32_t nt;                   // compact context rows per (p, c, m)
    uint32_t lbase[PR_RMAX + 2];   // first compact row of the q-node contexts; lbase[r + 1] = nt
    const uint32_t* map;           // [g.nt] full row -> compact row (0xffffffff: unused)
    const CtxRecipe* recipes;      // [nt] compact, parent_row compact
    const uint64_t* pair_tab;      // [K^(K-1)] pair_entry per tree index (context row compact), NULL: cut in the kernel
    // the same entries ordered by inside row (counting sort): the lanes of a warp then share their inside row, whose
    // loads become warp-uniform (the pair kernel is bound by L1 wavefronts).  Used for whole index ranges.
    const uint8_t* mask_owner;     // [2^K] node-set sharding of split clusterings: the rank of every pack node set,
                                   // balanced by the number of trees per set (NULL: round robin, node_set_owner)
    const FastProgram* irr_progs;  // programs of the irregular trees, indexed by the low 42 bits of their pair-table
                                   // entry (scre field 0); NULL: decode / per-K program table
    const uint32_t* sorted_tree;   // [K^(K-1)] tree index, NULL: none
    const uint64_t* sorted_entry;  // [K^(K-1)]
};

struct FastTables {
    const float* D2;     // [P][C][M][Kmax][L] log2 of the binomial tables               (clustering.py:20-24)
    const float* SD2;    // same layout: SD[r][i] = e^p0 / L^2 * sum_{x>=i} d_r[x]       (root without finishing forest)
    const float* FG;     // message tables, rows of L: G(F) of every forest F with <= s nodes; rows 0..K-1 of a
                         // block are the leaf messages  log2cumsum(D) + p0 - log L       (clustering.py:377)
    const float* FSG;    // scanned messages SG(F) = log2cumsum(G(F)) + p0 - log L
    const float* FH;     // finishing tables FH[r][F][i] = 1 / L^2 * sum_j d_r[i+j] * sg_F[j], forests with <= sh nodes
    const float* T;      // context tables (pairs.cuh), rows of L
    const TableBase* tb; // [P * C]
    const int16_t* lut;  // forest shape LUT (FT_LUT_SIZE entries)
    const ForestRecipe* const* recipes;   // [KMAX + 1] device arrays, by cluster count
    const uint32_t* const* level_order;   // [KMAX + 1]: per cluster count, the rows of every level with the
                                          // convolution rows first (position base[n] + i -> row)
    const PairK* pk;     // [KMAX + 1]
    const int32_t* pcm_list;   // [n_pcm] the (p * C + c) * M + m of the tables this rank builds (sharding plan)
    int n_pcm;
    int grid_blocks;           // table kernels: row blocks per (p, c, m) of this launch
    int shard_rank, shard_world;
};

__device__ __forceinline__ int set_owner(const PairK* pk, uint32_t mask, int world) {
    return pk->mask_owner ? (int)pk->mask_owner[mask] : node_set_owner(mask, world);
}

// ----------------------------------------------------------------------------- warp ops
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ float log2addexp(float a, float b) {
    const float m = fmaxf(a, b), d = fminf(a, b) - m;      // d <= 0 (NaN only if both -inf)
    if (!(d > -40.0f)) return m;                           // also catches -inf / NaN
    return m + lg2f(1.0f + ex2f(d));
}

__device__ __forceinline__ void warp_load_vec(float* __restrict__ dst, const float* __restrict__ src, int lane) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + lane + 32 * t);
        reinterpret_cast<float4*>(dst)[lane + 32 * t] = v;
    }
    __syncwarp();
}

__device__ __forceinline__ void warp_add_table(float* __restrict__ v, const float* __restrict__ src, int lane) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(src) + lane + 32 * t);
        float4 x = reinterpret_cast<float4*>(v)[lane + 32 * t];
        x.x += d.x; x.y += d.y; x.z += d.z; x.w += d.w;
        reinterpret_cast<float4*>(v)[lane + 32 * t] = x;
    }
    __syncwarp();
}

// in place: v[x] = log2(sum_{i<=x} 2^v[i]) + add                (trap_int / my_convolve(prior, .))
__device__ inline void warp_log2cumsum(float* __restrict__ v, float add, int lane) {
    float x[16];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float4 q = reinterpret_cast<const float4*>(v)[4 * lane + t];
        x[4 * t] = q.x; x[4 * t + 1] = q.y; x[4 * t + 2] = q.z; x[4 * t + 3] = q.w;
    }
    float m = x[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) m = fmaxf(m, x[i]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += ex2f(x[i] - m);
    float tot = m + lg2f(s);                                  // this lane's chunk total
    float run = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, run, d);
        if (lane >= d) run = log2addexp(o, run);
    }
    float r = __shfl_up_sync(0xffffffffu, run, 1);            // exclusive carry
    if (lane == 0) r = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 16; ++i) { r = log2addexp(r, x[i]); x[i] = r + add; }
#pragma unroll
    for (int t = 0; t < 4; ++t)
        reinterpret_cast<float4*>(v)[4 * lane + t] = make_float4(x[4 * t], x[4 * t + 1], x[4 * t + 2], x[4 * t + 3]);
    __syncwarp();
}

// log2(sum_x 2^v[x]) for all lanes
__device__ inline float warp_log2sumexp(const float* __restrict__ v, int lane) {
    float x[16];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float4 q = reinterpret_cast<const float4*>(v)[lane + 32 * t];
        x[4 * t] = q.x; x[4 * t + 1] = q.y; x[4 * t + 2] = q.z; x[4 * t + 3] = q.w;
    }
    float m = x[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) m = fmaxf(m, x[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += ex2f(x[i] - m);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    return m + lg2f(s);
}

// log2(sum_x 2^(v[x] + tab[x])), v in shared memory, tab a table row in global memory
// (tab == nullptr: plain log2sumexp of v) -- one code path for every way a sample is finished
__device__ inline float warp_log2dot(const float* __restrict__ v, const float* __restrict__ tab, int lane) {
    float x[16];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float4 q = reinterpret_cast<const float4*>(v)[lane + 32 * t];
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tab) d = __ldg(reinterpret_cast<const float4*>(tab) + lane + 32 * t);
        x[4 * t] = q.x + d.x; x[4 * t + 1] = q.y + d.y; x[4 * t + 2] = q.z + d.z; x[4 * t + 3] = q.w + d.w;
    }
    float m = x[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) m = fmaxf(m, x[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += ex2f(x[i] - m);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    return m + lg2f(s);
}

// merge path of the slope sequences of two log-concave vectors: jstar[x] = the b-index
// of the maximising split of output x (x = i + j).
__device__ inline void warp_merge_path(const float* __restrict__ a, const float* __restrict__ b,
                                       uint16_t* __restrict__ jstar, int lane) {
    const int x0 = 16 * lane;
    int lo = 0, hi = x0;                                      // x0 <= 496 < L, so i in [0, x0]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float sa = a[mid + 1] - a[mid];                 // gain of taking a's next step
        const float sb = b[x0 - mid] - b[x0 - mid - 1];       // gain of b's last taken step
        if (sa >= sb) lo = mid + 1; else hi = mid;
    }
    // state per side: the next value and the slope towards it; one load, one subtraction and a
    // handful of predicated moves per step
    int j = x0 - lo;
    const float* pa2 = a + lo + 2;
    const float* pb2 = b + j + 2;
    float an = a[lo + 1], bn = b[j + 1];
    float da = an - a[lo], db = bn - b[j];
    uint32_t packed = 0;                                      // two 16-bit j* per word
#pragma unroll 2
    for (int t = 0; t < 16; ++t) {
        if (t & 1) reinterpret_cast<uint32_t*>(jstar)[(x0 + t) >> 1] = packed | ((uint32_t)j << 16);
        else packed = (uint32_t)j;
        // advance along the larger slope (ties and NaNs go to a); past the right edge both vectors
        // carry four -inf entries (warp_init_guards), so a side that has run out cannot advance
        const bool take_b = db > da;
        const float nv = *(take_b ? pb2 : pa2);
        const float d = nv - (take_b ? bn : an);
        if (take_b) { ++j; db = d; bn = nv; ++pb2; } else { da = d; an = nv; ++pa2; }
    }
    __syncwarp();
}

// Exact two-pass evaluation of the 32 outputs of one pass over the full range j <= x,
// written in place over a[x].  Only used when the windowed sum looks inconsistent (inputs
// that are not log-concave beyond rounding noise); makes no structural assumption.
static __device__ __noinline__ void warp_logconv_exact_pass(float* __restrict__ a, const float* __restrict__ b,
                                                            int pass, int lane) {
    const int x = 32 * pass + lane, xmax = 32 * pass + 31;
    float mx = -CUDART_INF_F;
    for (int j = 0; j <= xmax; ++j) mx = fmaxf(mx, a[x - j] + b[j]);      // a[<0] is the -inf guard
    float s = 0.f;
    for (int j = 0; j <= xmax; ++j) s += ex2f((a[x - j] - mx) + b[j]);
    __syncwarp();
    a[x] = (mx == -CUDART_INF_F) ? mx : mx + lg2f(s) - LOG2_L;
}

// acc += sum over n4 chunks of 4 consecutive j (starting at the 4-aligned j0) of
// 2^((a[x-j] - mx) + b[j]); pa = &a[x - j0], pb = &b[j0].  b is a broadcast 128-bit load, a a
// conflict-free scalar load per term; the two shifts and the accumulation are packed fp32x2
// adds (FADD2), the exponentials MUFU.EX2.
__device__ __forceinline__ void conv_accumulate(const float* __restrict__ pa, const float4* __restrict__ pb, int n4,
                                                const float2 nm2, float2& acc0, float2& acc1, float2& acc2,
                                                float2& acc3) {
    const float4* const pend = pb + (n4 & ~1);
#pragma unroll 1
    while (pb != pend) {
        const float4 bv = pb[0], bw = pb[1];
        pb += 2;
        float2 t0 = make_float2(pa[0], pa[-1]), t1 = make_float2(pa[-2], pa[-3]);
        float2 t2 = make_float2(pa[-4], pa[-5]), t3 = make_float2(pa[-6], pa[-7]);
        pa -= 8;
        t0 = __fadd2_rn(__fadd2_rn(t0, nm2), make_float2(bv.x, bv.y));
        t1 = __fadd2_rn(__fadd2_rn(t1, nm2), make_float2(bv.z, bv.w));
        t2 = __fadd2_rn(__fadd2_rn(t2, nm2), make_float2(bw.x, bw.y));
        t3 = __fadd2_rn(__fadd2_rn(t3, nm2), make_float2(bw.z, bw.w));
        acc0 = __fadd2_rn(acc0, make_float2(ex2f(t0.x), ex2f(t0.y)));
        acc1 = __fadd2_rn(acc1, make_float2(ex2f(t1.x), ex2f(t1.y)));
        acc2 = __fadd2_rn(acc2, make_float2(ex2f(t2.x), ex2f(t2.y)));
        acc3 = __fadd2_rn(acc3, make_float2(ex2f(t3.x), ex2f(t3.y)));
    }
    if (n4 & 1) {
        const float4 bv = pb[0];
        float2 t0 = make_float2(pa[0], pa[-1]), t1 = make_float2(pa[-2], pa[-3]);
        t0 = __fadd2_rn(__fadd2_rn(t0, nm2), make_float2(bv.x, bv.y));
        t1 = __fadd2_rn(__fadd2_rn(t1, nm2), make_float2(bv.z, bv.w));
        acc0 = __fadd2_rn(acc0, make_float2(ex2f(t0.x), ex2f(t0.y)));
        acc1 = __fadd2_rn(acc1, make_float2(ex2f(t1.x), ex2f(t1.y)));
    }
}

// One pass of the exponential path: outputs 32*pass .. 32*pass+31, lane = output, in place.  Used by the
// framed instantiation for the blocks that do not fit a frame; the exponential-only instantiations
// run conv_pass_pair (two passes per step).
__device__ __forceinline__ void conv_pass_mufu(float* __restrict__ a, const float* __restrict__ b,
                                               const uint16_t* __restrict__ jstar, int pass, int lane, int& hl, int& hr,
                                               int& chunks) {
    const int x0 = 32 * pass;
    const int x = x0 + lane;
    const int xmax = x0 + 31;
    const int js = jstar[x];
    // min / max of the lanes' j* in two warp reductions (CREDUX)
    const int jmin = (int)__reduce_min_sync(0xffffffffu, (unsigned)js);
    const int jmax = (int)__reduce_max_sync(0xffffffffu, (unsigned)js);
    const float mx = a[x - js] + b[js];
    const float2 nm2 = make_float2(-mx, -mx);
    float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0, acc2 = acc0, acc3 = acc0;
    // first guess of the common j-range [jl, je] (4-aligned start, whole chunks); j may run
    // up to 3 past xmax: b reads its trailing -inf pad there, a[x - j] its -inf guard
    int jl = max(jmin - hl, 0) & ~3;
    int n4 = (min(jmax + hr, xmax) - jl + 4) >> 2;
    int je = jl + 4 * n4 - 1;
    // One call site for the accumulation loop (code size): first the guessed range, then -- as long
    // as some lane's edge term is still above 2^-THETA of its maximum -- extensions to the left,
    // then to the right.  The terms of one output are concave in j, so sub-threshold edges bound
    // everything outside the range.
    int seg_j0 = jl, seg_m4 = n4;
    bool left_open = jl > 0;                                   // warp-uniform
#pragma unroll 1
    for (;;) {
        chunks += seg_m4;
        conv_accumulate(a + (x - seg_j0), reinterpret_cast<const float4*>(b + seg_j0), seg_m4, nm2, acc0, acc1, acc2, acc3);
        // both edge terms at once: most passes need no extension and leave after ONE vote.  Lanes with
        // x <= je read the -inf guard on the right, so they never ask for more there.
        const float er = (a[x - je] - mx) + b[je];
        const float el = (a[x - jl] - mx) + b[jl];
        const bool need_r = (je < xmax) && er >= -THETA;
        const bool need_l = left_open && el >= -THETA;
        if (!__any_sync(0xffffffffu, need_l || need_r)) break;
        if (left_open) {
            if (__any_sync(0xffffffffu, need_l)) {
                const int j2 = max(jl - max(hl, 8), 0) & ~3;
                seg_j0 = j2; seg_m4 = (jl - j2) >> 2;
                hl += jl - j2;
                jl = j2;
                left_open = jl > 0;
                continue;
            }
            left_open = false;                                 // the left edge is settled for good
        }
        seg_m4 = min((max(hr, 8) + 3) >> 2, ((xmax | 3) - je) >> 2);
        seg_j0 = je + 1;
        hr += 4 * seg_m4;
        je += 4 * seg_m4;
    }
    acc0 = __fadd2_rn(__fadd2_rn(acc0, acc1), __fadd2_rn(acc2, acc3));
    const float s = acc0.x + acc0.y;
    // the range always contains the maximising term (2^0 up to the rounding of the shift,
    // which is relative to |mx|); a sum far outside [1, 2*terms] means M_x was not the
    // maximum, i.e. the inputs were not log-concave
    const bool bad = !(s >= 0.7f && s <= 2048.f);
    if (__any_sync(0xffffffffu, bad)) {
        warp_logconv_exact_pass(a, b, pass, lane);
    } else {
        __syncwarp();                                         // every lane has finished reading a[x0..xmax]
        a[x] = mx + lg2f(s) - LOG2_L;
    }
    // let the windows shrink again (they are re-grown on demand)
    hl = max(4, hl - (hl >> 2));
    hr = max(4, hr - (hr >> 2));
}

// ---- two passes at once: each half-warp takes one 32-output pass, every lane two adjacent outputs ----
// acc += sum over n4 chunks of 4 consecutive j (from the 4-aligned j0) of 2^((a[x-j] - mx0) + b[j]) in .x
// and 2^((a[x+1-j] - mx1) + b[j]) in .y; x even, pa = &a[x - j0] (8-byte aligned), pb = &b[j0].  The two
// outputs of a lane share every loaded element: 2 LDS.64 + 1 LDS.128 per 8 terms instead of 8 LDS.32 +
// 2 LDS.128.  n4 may differ between the half-warps (each runs its own window).
__device__ __forceinline__ void conv_accumulate_pair(const float* __restrict__ pa, const float4* __restrict__ pb,
                                                     int n4, const float2 nm, float2& acc0, float2& acc1,
                                                     float2& acc2, float2& acc3) {
    if (n4 <= 0) return;
    float2 q0 = *reinterpret_cast<const float2*>(pa);         // (a[p], a[p+1]), p = x - j0
    const float4* const pend = pb + (n4 & ~1);
#pragma unroll 1
    while (pb != pend) {
        const float4 bv = pb[0], bw = pb[1];
        pb += 2;
        const float2 q1 = *reinterpret_cast<const float2*>(pa - 2), q2 = *reinterpret_cast<const float2*>(pa - 4);
        const float2 q3 = *reinterpret_cast<const float2*>(pa - 6), q4 = *reinterpret_cast<const float2*>(pa - 8);
        pa -= 8;
        // j0 + k, k = 0..7: the pair (a[p-k], a[p-k+1]) is q_{k/2} for even k, (q_{(k+1)/2}.y, q_{(k-1)/2}.x) for odd k
        float2 t0 = __fadd2_rn(__fadd2_rn(q0, nm), make_float2(bv.x, bv.x));
        float2 t1 = __fadd2_rn(__fadd2_rn(make_float2(q1.y, q0.x), nm), make_float2(bv.y, bv.y));
        float2 t2 = __fadd2_rn(__fadd2_rn(q1, nm), make_float2(bv.z, bv.z));
        float2 t3 = __fadd2_rn(__fadd2_rn(make_float2(q2.y, q1.x), nm), make_float2(bv.w, bv.w));
        acc0 = __fadd2_rn(acc0, make_float2(ex2f(t0.x), ex2f(t0.y)));
        acc1 = __fadd2_rn(acc1, make_float2(ex2f(t1.x), ex2f(t1.y)));
        acc2 = __fadd2_rn(acc2, make_float2(ex2f(t2.x), ex2f(t2.y)));
        acc3 = __fadd2_rn(acc3, make_float2(ex2f(t3.x), ex2f(t3.y)));
        t0 = __fadd2_rn(__fadd2_rn(q2, nm), make_float2(bw.x, bw.x));
        t1 = __fadd2_rn(__fadd2_rn(make_float2(q3.y, q2.x), nm), make_float2(bw.y, bw.y));
        t2 = __fadd2_rn(__fadd2_rn(q3, nm), make_float2(bw.z, bw.z));
        t3 = __fadd2_rn(__fadd2_rn(make_float2(q4.y, q3.x), nm), make_float2(bw.w, bw.w));
        acc0 = __fadd2_rn(acc0, make_float2(ex2f(t0.x), ex2f(t0.y)));
        acc1 = __fadd2_rn(acc1, make_float2(ex2f(t1.x), ex2f(t1.y)));
        acc2 = __fadd2_rn(acc2, make_float2(ex2f(t2.x), ex2f(t2.y)));
        acc3 = __fadd2_rn(acc3, make_float2(ex2f(t3.x), ex2f(t3.y)));
        q0 = q4;
    }
    if (n4 & 1) {
        const float4 bv = pb[0];
        const float2 q1 = *reinterpret_cast<const float2*>(pa - 2), q2 = *reinterpret_cast<const float2*>(pa - 4);
        const float2 t0 = __fadd2_rn(__fadd2_rn(q0, nm), make_float2(bv.x, bv.x));
        const float2 t1 = __fadd2_rn(__fadd2_rn(make_float2(q1.y, q0.x), nm), make_float2(bv.y, bv.y));
        const float2 t2 = __fadd2_rn(__fadd2_rn(q1, nm), make_float2(bv.z, bv.z));
        const float2 t3 = __fadd2_rn(__fadd2_rn(make_float2(q2.y, q1.x), nm), make_float2(bv.w, bv.w));
        acc0 = __fadd2_rn(acc0, make_float2(ex2f(t0.x), ex2f(t0.y)));
        acc1 = __fadd2_rn(acc1, make_float2(ex2f(t1.x), ex2f(t1.y)));
        acc2 = __fadd2_rn(acc2, make_float2(ex2f(t2.x), ex2f(t2.y)));
        acc3 = __fadd2_rn(acc3, make_float2(ex2f(t3.x), ex2f(t3.y)));
    }
}

// Passes 2q+1 (lanes 16..31) and 2q (lanes 0..15) together: outputs 64q .. 64q+63, lane (h, m) owns
// x = 64q + 32h + 2m and x + 1.  Each half-warp keeps its own j-window (jl, je and the extension state
// are uniform inside a half), so nothing is summed that the 32-output pass would not sum; the
// bookkeeping instructions are shared by both passes.  In place: both halves have finished reading
// before anything is written.
__device__ __forceinline__ void conv_pass_pair(float* __restrict__ a, const float* __restrict__ b,
                                               const uint16_t* __restrict__ jstar, int q, int lane, int& hl, int& hr,
                                               int& chunks) {
    const int x0 = 64 * q + (lane & 16) * 2;                  // first output of this half's pass
    const int x = x0 + 2 * (lane & 15);
    const int xmax = x0 + 31;
    const uint32_t jpair = reinterpret_cast<const uint32_t*>(jstar)[x >> 1];
    const int js0 = (int)(jpair & 0xffffu), js1 = (int)(jpair >> 16);
    // j* is non-decreasing inside each 16-output walk of the merge path
    const int jmin = min((int)jstar[x0], (int)jstar[x0 + 16]);
    const int jmax = max((int)jstar[x0 + 15], (int)jstar[xmax]);
    const float mx0 = a[x - js0] + b[js0];
    const float mx1 = a[x + 1 - js1] + b[js1];
    const float2 nm = make_float2(-mx0, -mx1);
    float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0, acc2 = acc0, acc3 = acc0;
    int jl = max(jmin - hl, 0) & ~3;
    int n4 = (min(jmax + hr, xmax) - jl + 4) >> 2;
    int je = jl + 4 * n4 - 1;
    int seg_j0 = jl, seg_m4 = n4;
    bool left_open = jl > 0;                                   // uniform inside a half
    const unsigned half_mask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;
    int grow_l = 0, grow_r = 0;
#pragma unroll 1
    for (;;) {
        chunks += seg_m4;
        conv_accumulate_pair(a + (x - seg_j0), reinterpret_cast<const float4*>(b + seg_j0), seg_m4, nm, acc0, acc1, acc2, acc3);
        // edge terms of both outputs on both sides; lanes whose index runs below 0 read the -inf guard
        const float2 al = *reinterpret_cast<const float2*>(a + (x - jl));             // x - jl is even
        const float bl = b[jl], br = b[je];
        const float ar0 = a[x - je], ar1 = a[x + 1 - je];
        const bool need_l = left_open && (((al.x - mx0) + bl) >= -THETA || ((al.y - mx1) + bl) >= -THETA);
        const bool need_r = (je < xmax) && (((ar0 - mx0) + br) >= -THETA || ((ar1 - mx1) + br) >= -THETA);
        const unsigned vote_l = __ballot_sync(0xffffffffu, need_l), vote_r = __ballot_sync(0xffffffffu, need_r);
        if ((vote_l | vote_r) == 0) break;
        seg_m4 = 0;                                            // a half that is done idles through the other's extension
        if (left_open) {
            if (vote_l & half_mask) {
                const int j2 = max(jl - max(hl + grow_l, 8), 0) & ~3;
                seg_j0 = j2; seg_m4 = (jl - j2) >> 2;
                grow_l += jl - j2;
                jl = j2;
                left_open = jl > 0;
                continue;
            }
            left_open = false;                                 // this half's left edge is settled for good
        }
        if (vote_r & half_mask) {
            seg_m4 = min((max(hr + grow_r, 8) + 3) >> 2, ((xmax | 3) - je) >> 2);
            seg_j0 = je + 1;
            grow_r += 4 * seg_m4;
            je += 4 * seg_m4;
        }
    }
    acc0 = __fadd2_rn(__fadd2_rn(acc0, acc1), __fadd2_rn(acc2, acc3));
    const bool bad = !(acc0.x >= 0.7f && acc0.x <= 2048.f && acc0.y >= 0.7f && acc0.y <= 2048.f);
    if (__any_sync(0xffffffffu, bad)) {
        warp_logconv_exact_pass(a, b, 2 * q + 1, lane);
        warp_logconv_exact_pass(a, b, 2 * q, lane);
    } else {
        __syncwarp();                                         // every lane has finished reading a[64q .. 64q+63]
        *reinterpret_cast<float2*>(a + x) = make_float2(mx0 + lg2f(acc0.x) - LOG2_L, mx1 + lg2f(acc0.y) - LOG2_L);
    }
    // window half-widths: grown by the larger extension of the two halves, then allowed to shrink again
    grow_l = max(grow_l, __shfl_xor_sync(0xffffffffu, grow_l, 16));
    grow_r = max(grow_r, __shfl_xor_sync(0xffffffffu, grow_r, 16));
    hl += grow_l; hr += grow_r;
    hl = max(4, hl - (hl >> 2));
    hr = max(4, hr - (hr >> 2));
}

// One block of the LINEAR-DOMAIN ("framed") path: outputs 64*q .. 64*q+63, two per lane, in place.
// Inside a block M_x is close to linear in x.  After removing a common tilt s (a multiple of 1/64, so
// s * index is exact) and the centre output's maximising factors (ca, cb), every factor that matters
// fits fp32:  fa[e] = 2^(a[i] - ca - s (i - i0)),  fb[e] = 2^(b[j] - cb - s (j - j0)), and
// sum_j fa[x-j] fb[j] = 2^-(ca + cb + s (x - xc)) sum_j a[x-j] b[j].  The exponentials are paid once
// per ELEMENT (~9 per lane and segment) instead of once per TERM, the terms are plain FFMAs with
// a register-resident sliding window (2 outputs x 4 terms per 2 LDS.64 + 1 broadcast LDS.128).
// Returns false (warp-uniform, nothing written) when the block does not fit a frame: M_x bends by
// more than 2^FR_DEV, a factor exceeds 2^FR_MAXR, or the result fails the sanity check; the caller
// then runs the exponential path on the two halves.
__device__ __forceinline__ bool conv_block_framed(float* __restrict__ a, const float* __restrict__ b,
                                                  const uint16_t* __restrict__ jstar, float* __restrict__ fa,
                                                  float* __restrict__ fb, int q, int lane, int& hl, int& hr,
                                                  int& chunks) {
    const int x0 = FR_B * q, xlo = x0 + 2 * lane, xmax = x0 + FR_B - 1;
    const uint32_t jpair = reinterpret_cast<const uint32_t*>(jstar)[xlo >> 1];
    const int js0 = (int)(jpair & 0xffffu), js1 = (int)(jpair >> 16);
    const float mx0 = a[xlo - js0] + b[js0];
    const float mx1 = a[xlo + 1 - js1] + b[js1];
    const float mx_c = __shfl_sync(0xffffffffu, mx0, 16);
    const int j0 = __shfl_sync(0xffffffffu, js0, 16);
    const int i0 = x0 + 32 - j0;
    const float s = rintf((__shfl_sync(0xffffffffu, mx1, 31) - __shfl_sync(0xffffffffu, mx0, 0)) * (64.0f / 63.0f)) *
                    (1.0f / 64.0f);
    const float dev0 = mx0 - fmaf(s, (float)(2 * lane - 32), mx_c);      // M_x minus its linear model
    const float dev1 = mx1 - fmaf(s, (float)(2 * lane - 31), mx_c);
    if (__any_sync(0xffffffffu, !(fabsf(dev0) <= FR_DEV && fabsf(dev1) <= FR_DEV)) || !(fabsf(s) < 256.f)) return false;
    const float ca = a[i0], cb = b[j0];                       // ca + cb == mx_c: the centre's best term is 2^0
    const int jmin = __reduce_min_sync(0xffffffffu, min(js0, js1));
    const int jmax = __reduce_max_sync(0xffffffffu, max(js0, js1));
    int jl = max(jmin - hl, 0) & ~3;
    int n4 = (min(jmax + hr, xmax) - jl + 4) >> 2;
    int je = jl + 4 * n4 - 1;
    float f00 = 0.f, f01 = 0.f, f10 = 0.f, f11 = 0.f;
    int my_chunks = 0;
    int seg_j0 = jl, seg_m4 = n4, phase = 0;
#pragma unroll 1
    for (;;) {
        my_chunks += 2 * seg_m4;
#pragma unroll 1
        for (int o = 0; o < seg_m4; o += FR_SEG / 4) {
            const int mm = min(FR_SEG / 4, seg_m4 - o);
            const int j0s = seg_j0 + 4 * o, nb = 4 * mm;
            const int i_lo = x0 - (j0s + nb - 1);
            bool over = false;
#pragma unroll 1
            for (int k = 0; k < FR_SEG / 32; ++k) {
                const int e = lane + 32 * k;
                if (e < nb) {
                    const int j = j0s + e;
                    const float r = fmaf(-s, (float)(j - j0), b[j] - cb);
                    over |= r > FR_MAXR;
                    fb[e] = ex2f(r);
                }
            }
#pragma unroll 1
            for (int k = 0; k < FR_A / 32; ++k) {
                const int e = lane + 32 * k;
                if (e < nb + FR_B - 1) {
                    const int i = i_lo + e;                   // may be below -GUARD: a 64-wide block reaches 66 back
                    const float r = fmaf(-s, (float)(i - i0), a[max(i, -GUARD)] - ca);
                    over |= r > FR_MAXR;
                    fa[e] = ex2f(r);
                }
            }
            if (__any_sync(0xffffffffu, over)) return false;
            __syncwarp();
            // sliding window: output lo at j uses fa[E - (j - j0s)], E = 2 lane + nb - 1; output hi the next one
            const float* pw = fa + 2 * lane + nb - 4;         // &fa[E - 3], even index
            float w1 = pw[1], w2 = pw[2], w3 = pw[3], w4 = pw[4], w0 = pw[0];
            const float4* pB = reinterpret_cast<const float4*>(fb);
#pragma unroll 2
            for (int c = 0; c < mm; ++c) {
                const float4 bv = pB[c];
                // lo: fa[E]*b0 + fa[E-1]*b1 + fa[E-2]*b2 + fa[E-3]*b3 ; hi: fa[E+1]*b0 + fa[E]*b1 + ...
                f00 = fmaf(w3, bv.x, f00); f10 = fmaf(w4, bv.x, f10);
                f01 = fmaf(w2, bv.y, f01); f11 = fmaf(w3, bv.y, f11);
                f00 = fmaf(w1, bv.z, f00); f10 = fmaf(w2, bv.z, f10);
                f01 = fmaf(w0, bv.w, f01); f11 = fmaf(w1, bv.w, f11);
                pw -= 4;
                w4 = w0;
                const float2 u = *reinterpret_cast<const float2*>(pw);        // fa[E-7], fa[E-6]
                const float2 v = *reinterpret_cast<const float2*>(pw + 2);    // fa[E-5], fa[E-4]
                w0 = u.x; w1 = u.y; w2 = v.x; w3 = v.y;
            }
            __syncwarp();                                     // scratch is reused by the next segment
        }
        if (phase == 0) {
            const bool need = (jl > 0) && ((((a[xlo - jl] - mx0) + b[jl]) >= -THETA) ||
                                           (((a[xlo + 1 - jl] - mx1) + b[jl]) >= -THETA));   // jl <= j* <= xlo
            if (__any_sync(0xffffffffu, need)) {
                const int j2 = max(jl - max(hl, 8), 0) & ~3;
                seg_j0 = j2; seg_m4 = (jl - j2) >> 2;
                hl += jl - j2;
                jl = j2;
                continue;
            }
            phase = 1;
        }
        // lanes whose outputs end before je read the -inf guard here, so they never ask for more
        const bool need = (((a[max(xlo - je, -GUARD)] - mx0) + b[je]) >= -THETA) ||
                          (((a[max(xlo + 1 - je, -GUARD)] - mx1) + b[je]) >= -THETA);
        if (!__any_sync(0xffffffffu, need) || je >= xmax) break;
        seg_m4 = min((max(hr, 8) + 3) >> 2, ((xmax | 3) - je) >> 2);
        seg_j0 = je + 1;
        hr += 4 * seg_m4;
        je += 4 * seg_m4;
    }
    // framed sums are relative to 2^(mx_c + s (x - xc)); bring them to 2^mx
    const float s0 = (f00 + f01) * ex2f(-dev0), s1 = (f10 + f11) * ex2f(-dev1);
    const bool bad = !(s0 >= 0.7f && s0 <= 2048.f && s1 >= 0.7f && s1 <= 2048.f);
    if (__any_sync(0xffffffffu, bad)) return false;
    __syncwarp();                                             // every lane has finished reading a[x0..xmax]
    *reinterpret_cast<float2*>(a + xlo) = make_float2(mx0 + lg2f(s0) - LOG2_L, mx1 + lg2f(s1) - LOG2_L);
    chunks += my_chunks;
    hl = max(4, hl - (hl >> 2));
    hr = max(4, hr - (hr >> 2));
    return true;
}

// IN PLACE: a[x] <- log2( sum_{j<=x} 2^(a[x-j] + b[j]) ) - log2 L          (clustering.py:365-370)
// a, b: distinct shared-memory vectors, each with GUARD -inf floats in front and 4 behind.
// Blocks run from the highest outputs down: output x only needs a[i], i <= x, so finished
// outputs can overwrite a[x] once their block has read everything it needs.
// fa / fb: per-warp scratch (FR_A / FR_SEG floats) of the framed path; FRAMED = false compiles
// the exponential path only.  Returns the number of 4-term chunks executed per lane.
template <bool FRAMED>
__device__ inline int warp_logconv(float* __restrict__ a, const float* __restrict__ b,
                                   uint16_t* __restrict__ jstar, float* __restrict__ fa, float* __restrict__ fb,
                                   int lane) {
    warp_merge_path(a, b, jstar, lane);
    int chunks = 0;
    int hl = 8, hr = 8;                                       // window half-widths, adapted block to block
    // keep the two vector bases in registers across the passes (ptxas otherwise rebuilds them from
    // the warp base and the stack position several times per pass)
    unsigned sa = (unsigned)__cvta_generic_to_shared(a), sb = (unsigned)__cvta_generic_to_shared(b);
    asm volatile("" : "+r"(sa), "+r"(sb));
    a = reinterpret_cast<float*>(__cvta_shared_to_generic(sa));
    b = reinterpret_cast<const float*>(__cvta_shared_to_generic(sb));
    if (FRAMED) {
#pragma unroll 1
        for (int q = L / FR_B - 1; q >= 0; --q) {
            if (conv_block_framed(a, b, jstar, fa, fb, q, lane, hl, hr, chunks)) continue;
#pragma unroll 1
            for (int h = 1; h >= 0; --h) conv_pass_mufu(a, b, jstar, 2 * q + h, lane, hl, hr, chunks);
        }
    } else {
#pragma unroll 1
        for (int q = L / 64 - 1; q >= 0; --q) conv_pass_pair(a, b, jstar, q, lane, hl, hr, chunks);
        chunks += __shfl_xor_sync(0xffffffffu, chunks, 16);   // 4-term chunks of both passes (per output lane)
    }
    __syncwarp();
    return chunks;
}

}  // namespace bamse
