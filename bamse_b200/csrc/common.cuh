// Shared definitions of libbamse_cuda (sm_100a).  See include/bamse_cuda.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <math_constants.h>

#include "../../include/bamse_cuda.h"

namespace bamse {

constexpr int L = BAMSE_GRID;          // prevalence grid points (clustering.py:20)
constexpr int KMAX = BAMSE_KMAX;
constexpr int MAX_OPS = 2 * KMAX + 2;  // evaluation program length bound
constexpr int MAX_STACK = KMAX / 2 + 3;  // live vectors in reference child order (+1 conv scratch)

// Device-resident problem (uploaded by bamse_upload_problem).
struct ProblemView {
    int P, C, M, Kmax;
    const int32_t* K_of_c;    // [C]
    const double* p0;         // [C]   -lgamma(K)/K                      (clustering.py:29)
    const double* D64;        // [P][C][M][Kmax][L] log tables, fp64      (clustering.py:20-24)
    const float* D32;         // same, rounded to fp32 (log-domain fp32 kernel)
    __host__ __device__ size_t table_offset(int p, int c, int m) const {
        return (((size_t)p * C + c) * M + m) * (size_t)Kmax * L;
    }
};

// Evaluation program of one tree: a post-order walk in the reference's child order
// (children by increasing index, clustering.py:372-379 + tree_tools.py:105-107).
enum : int8_t {
    OP_LEAF = 0,   // push E(leaf v) = logcumsumexp(D[v]) + p0 - log L      (clustering.py:377)
    OP_SCAN = 1,   // top = logcumsumexp(top) + p0 - log L   (first child: my_convolve(prior, E))
    OP_CONV = 2,   // b = pop; top = my_convolve(top, b)                    (clustering.py:379)
    OP_ADDD = 3,   // top += D[v]                                           (clustering.py:379)
};

struct TreeProgram {
    int8_t n_ops;
    int8_t k;              // nodes in this tree
    int8_t n_leaves;
    int8_t pad;            // fast path: how the per-sample scalar is finished (FIN_*)
    int8_t op[MAX_OPS];
    int8_t arg[MAX_OPS];   // table row (cluster id) for OP_LEAF / OP_ADDD
    int8_t fin_r, fin_c;   // fast path: root row / leaf-child row of the finishing table
    int32_t scre;          // canonizeF 'scre'                              (tree_tools.py:47-63)
};

__host__ __device__ inline uint64_t ipow_u64(int base, int e) {
    uint64_t r = 1;
    for (int i = 0; i < e; ++i) r *= (uint64_t)base;
    return r;
}

}  // namespace bamse

// -DBAMSE_DEBUG_BOUNDS (make DEBUG=1 -> libbamse_cuda_dbg.so): every table-row / list index the kernels compute is
// checked against its extent and a violation traps (the launch fails with "unspecified launch failure", the tests
// fail loudly).  compute-sanitizer is closed on the GPU pool this was developed on; running the GPU test suite on
// the checked build (profiles/bounds_check.sh) is the substitute for its memcheck pass.
#ifdef BAMSE_DEBUG_BOUNDS
#define BAMSE_BOUND(cond) do { if (!(cond)) asm volatile("trap;"); } while (0)
#else
#define BAMSE_BOUND(cond) do { } while (0)
#endif

#define BAMSE_CUDA_TRY(ctx, expr)                                                        \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            (ctx)->fail(BAMSE_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                             \
            return BAMSE_ECUDA;                                                          \
        }                                                                                \
    } while (0)
