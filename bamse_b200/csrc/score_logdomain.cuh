// Log-domain scorer (template on the arithmetic type).
//   T = double : the fp64 CHECK MODE -- follows the reference operation by operation
//                (two-pass max-shifted log-sum-exp per convolution output).
//   T = float  : same structure in fp32 with MUFU ex2/lg2 (XU-pipe bound); kept as the
//                robust any-dynamic-range fp32 path.
// One CTA evaluates one tree at a time over all M samples; 256 threads, each owning the
// output pair (x, L-1-x) of a convolution so every thread sums exactly L+1 terms.
#pragma once
#include "common.cuh"
#include "tree_decode.cuh"

namespace bamse {

constexpr int LD_THREADS = 256;

template <typename T> struct LMath;
template <> struct LMath<double> {
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double log1p_(double x) { return log1p(x); }
    static __device__ __forceinline__ double ninf() { return -CUDART_INF; }
};
template <> struct LMath<float> {
    static __device__ __forceinline__ float exp_(float x) { return __expf(x); }
    static __device__ __forceinline__ float log_(float x) { return __logf(x); }
    static __device__ __forceinline__ float log1p_(float x) { return __logf(1.0f + x); }
    static __device__ __forceinline__ float ninf() { return -CUDART_INF_F; }
};

template <typename T>
__device__ __forceinline__ T logaddexp_(T a, T b) {          // np.logaddexp (clustering.py:362)
    const T m = a > b ? a : b;
    const T d = a > b ? b - a : a - b;                        // <= 0
    if (m == LMath<T>::ninf()) return m;
    return m + LMath<T>::log1p_(LMath<T>::exp_(d));
}

template <typename T>
__device__ __forceinline__ T shfl_up_(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
template <typename T>
__device__ __forceinline__ T shfl_xor_(T v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }

// In-place inclusive log-cumsum of v[0..L) plus `add` (block-wide, LD_THREADS threads).
// Restates trap_int / my_convolve(prior, .) (clustering.py:359-363, 377).
template <typename T>
__device__ void block_logcumsum(T* v, T add, T* warp_tot /* [LD_THREADS/32] */) {
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    T x0 = v[2 * t], x1 = v[2 * t + 1];
    x1 = logaddexp_(x0, x1);
    T run = x1;                                               // inclusive over this thread's pair
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up_(run, d);
        if (lane >= d) run = logaddexp_(o, run);
    }
    if (lane == 31) warp_tot[w] = run;
    __syncthreads();
    T prefix = LMath<T>::ninf();                              // exclusive prefix of preceding warps
    for (int i = 0; i < w; ++i) prefix = logaddexp_(prefix, warp_tot[i]);
    T excl = shfl_up_(run, 1);                                // exclusive within warp
    if (lane == 0) excl = LMath<T>::ninf();
    excl = logaddexp_(prefix, excl);
    v[2 * t] = logaddexp_(excl, x0) + add;
    v[2 * t + 1] = logaddexp_(excl, x1) + add;
    __syncthreads();
}

// out[x] = logsumexp_{j<=x}(b[j] + a[x-j]) - log L        (clustering.py:365-370)
template <typename T>
__device__ void block_logconv(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                              T log_len) {
    const int t = threadIdx.x;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int x = half == 0 ? t : L - 1 - t;
        T mx = LMath<T>::ninf();
        for (int j = 0; j <= x; ++j) {
            const T v = b[j] + a[x - j];
            mx = v > mx ? v : mx;
        }
        T s = 0;
        if (mx != LMath<T>::ninf())
            for (int j = 0; j <= x; ++j) s += LMath<T>::exp_(b[j] + a[x - j] - mx);
        out[x] = (mx == LMath<T>::ninf()) ? mx : (LMath<T>::log_(s) + mx - log_len);
    }
    __syncthreads();
}

// logsumexp(v) - log L   (trap_int(...)[-1], clustering.py:385); result valid in all threads.
template <typename T>
__device__ T block_logsumexp(const T* v, T log_len, T* red /* [LD_THREADS/32] */) {
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    T a = v[t], b = v[t + LD_THREADS];
    T mx = a > b ? a : b;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { T o = shfl_xor_(mx, d); mx = o > mx ? o : mx; }
    if (lane == 0) red[w] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < LD_THREADS / 32; ++i) mx = red[i] > mx ? red[i] : mx;
    __syncthreads();
    if (mx == LMath<T>::ninf()) return mx;
    T s = LMath<T>::exp_(a - mx) + LMath<T>::exp_(b - mx);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += shfl_xor_(s, d);
    if (lane == 0) red[w] = s;
    __syncthreads();
    s = red[0];
    for (int i = 1; i < LD_THREADS / 32; ++i) s += red[i];
    __syncthreads();
    return LMath<T>::log_(s) + mx - log_len;
}

template <typename T> struct TableSel;
template <> struct TableSel<double> { static __device__ const double* get(const ProblemView& pv) { return pv.D64; } };
template <> struct TableSel<float> { static __device__ const float* get(const ProblemView& pv) { return pv.D32; } };

// Scores one tree (program in shared memory) over all samples of (p, c).  All threads of
// the CTA call this; the fp64 sum over samples (np.sum, clustering.py:89) is returned
// in every thread.
template <typename T>
__device__ double eval_tree_logdomain(const ProblemView& pv, int p, int c, const TreeProgram* prog,
                                      T* pool /* [MAX_STACK][L] */, T* scratch /* [LD_THREADS/32] */,
                                      int m_begin = 0, int m_end = -1, double cut = -HUGE_VAL,
                                      unsigned* n_done = nullptr) {
    const int t = threadIdx.x;
    const T log_len = (T)6.2383246250395077847550890931236;  // log(512)
    const T add = (T)pv.p0[c] - log_len;
    const T* table = TableSel<T>::get(pv);
    double total = 0.0;
    if (m_end < 0) m_end = pv.M;
    for (int m = m_begin; m < m_end; ++m) {
        const T* D = table + pv.table_offset(p, c, m);
        // slot indirection: stack position i lives in pool slot (slots >> 4i) & 15
        uint64_t slots = 0x876543210ull;
        int sp = 0;
        const int n_ops = prog->n_ops;
        for (int i = 0; i < n_ops; ++i) {
            const int op = prog->op[i];
            if (op == OP_LEAF) {
                T* v = pool + ((slots >> (4 * sp)) & 15) * L;
                const T* d = D + (size_t)prog->arg[i] * L;
                v[t] = d[t]; v[t + LD_THREADS] = d[t + LD_THREADS];
                __syncthreads();
                block_logcumsum(v, add, scratch);
                ++sp;
            } else if (op == OP_SCAN) {
                T* v = pool + ((slots >> (4 * (sp - 1))) & 15) * L;
                block_logcumsum(v, add, scratch);
            } else if (op == OP_CONV) {
                const int sa = (int)((slots >> (4 * (sp - 2))) & 15);
                const int sb = (int)((slots >> (4 * (sp - 1))) & 15);
                const int so = (int)((slots >> (4 * sp)) & 15);
                block_logconv(pool + sa * L, pool + sb * L, pool + so * L, log_len);
                // result becomes stack[sp-2]; old a-slot becomes the free scratch at [sp]
                slots &= ~((0xFull << (4 * (sp - 2))) | (0xFull << (4 * sp)));
                slots |= ((uint64_t)so << (4 * (sp - 2))) | ((uint64_t)sa << (4 * sp));
                --sp;
            } else {  // OP_ADDD
                T* v = pool + ((slots >> (4 * (sp - 1))) & 15) * L;
                const T* d = D + (size_t)prog->arg[i] * L;
                v[t] += d[t]; v[t + LD_THREADS] += d[t + LD_THREADS];
                __syncthreads();
            }
        }
        const T sm = block_logsumexp(pool + (slots & 15) * L, log_len, scratch);
        total += (double)sm;
        if (n_done) ++*n_done;
        // exact early exit (BAMSE_OPT_EARLY_EXIT): every per-sample score is <= 0, so a partial sum below
        // the cut cannot recover; block-uniform (every thread holds the same total)
        if (total < cut) return -CUDART_INF;
    }
    return total;
}

}  // namespace bamse
