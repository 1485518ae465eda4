/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of ULTRA's generalized relational SpMM (rspmm).
 *
 * This file is the parity oracle for the sm_100a kernels in ultra_b200/csrc/.  It is NOT a product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.
 *
 * It restates, in plain scalar C, the algorithm of the reference CPU kernel:
 *   forward  loop : /root/reference/ultra/rspmm/source/rspmm.cpp:50-75
 *   backward loop : /root/reference/ultra/rspmm/source/rspmm.cpp:77-119
 *   functors      : /root/reference/ultra/rspmm/source/operator.cuh:13-80
 *   CSR pointer   : /root/reference/ultra/rspmm/source/rspmm.cpp:40-48   (ind2ptr)
 *   sort key      : /root/reference/ultra/rspmm/rspmm.py:174-178         (row-major (node_in, node_out))
 *
 * Pinning: the reference ships no golden vectors.  This restatement is pinned against the
 * reference's own CPU extension compiled from /root/reference (oracle/_ref, see oracle/build_ref.py) and
 * against the hand-computed case of SURVEY.md section 8c; the resulting fixtures live in tests/golden/.
 *
 * Semantics kept from the reference:
 *   - rows are edge_index[0], gathered columns are edge_index[1]            (rspmm.cu:238-240)
 *   - y = w * Bin(rel, in), edges of a row folded sequentially in CSR order (rspmm.cpp:62-71)
 *   - empty rows yield the Nary identity: 0 / +MAX / -MAX                   (operator.cuh:53,66,79)
 *   - min/max backward sends gradient to EVERY edge with y == out           (operator.cuh:62-64,75-77)
 *   - weight_grad[e] = sum_d og * dout_dy * x  (always computed on CPU)     (rspmm.cpp:104,115)
 * Compiled without -ffast-math so that the evaluation order below is the one executed.
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { SUM_ADD = 0, SUM_MIN = 1, SUM_MAX = 2 };
enum { MUL_MUL = 0, MUL_ADD = 1 };

/* rspmm.cpp:40-48 -- row pointer (N+1 entries here; the reference keeps N and uses nnz for the last end) */
static void ind2ptr(const int64_t *row_ind, int64_t nnz, int64_t num_row, int64_t *row_ptr) {
    memset(row_ptr, 0, sizeof(int64_t) * (size_t)(num_row + 1));
    for (int64_t e = 0; e < nnz; e++) row_ptr[row_ind[e] + 1]++;
    for (int64_t r = 0; r < num_row; r++) row_ptr[r + 1] += row_ptr[r];
}

#define DEFINE_ORACLE(T, SUFFIX, TMAX)                                                                         \
    static inline T bin_fwd_##SUFFIX(int mul, T rel, T in) { return mul == MUL_MUL ? rel * in : rel + in; }    \
    static inline T bin_dlhs_##SUFFIX(int mul, T rel, T in) { (void)rel; return mul == MUL_MUL ? in : (T)1; }  \
    static inline T bin_drhs_##SUFFIX(int mul, T rel, T in) { (void)in; return mul == MUL_MUL ? rel : (T)1; }  \
    static inline T nary_zero_##SUFFIX(int sum) {                                                              \
        return sum == SUM_ADD ? (T)0 : (sum == SUM_MIN ? (T)TMAX : (T)(-TMAX));                                \
    }                                                                                                          \
    static inline T nary_fwd_##SUFFIX(int sum, T r, T x) {                                                     \
        if (sum == SUM_ADD) return r + x;                                                                      \
        if (sum == SUM_MIN) return r < x ? r : x;                                                              \
        return r > x ? r : x;                                                                                  \
    }                                                                                                          \
    static inline T nary_bwd_##SUFFIX(int sum, T r, T x) {                                                     \
        if (sum == SUM_ADD) return (T)1;                                                                       \
        return r == x ? (T)1 : (T)0;                                                                           \
    }                                                                                                          \
    /* rspmm.cpp:50-75.  edge_index is (2, nnz) row-major, already sorted by (row, col). */                    \
    int rspmm_oracle_forward_##SUFFIX(int sum, int mul, const int64_t *edge_index, const int64_t *edge_type,   \
                                      const T *edge_weight, const T *relation, const T *input, T *output,      \
                                      int64_t num_row, int64_t nnz, int64_t dim) {                             \
        const int64_t *row_ind = edge_index, *col_ind = edge_index + nnz;                                      \
        int64_t *row_ptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(num_row + 1));                         \
        if (!row_ptr) return 1;                                                                                \
        ind2ptr(row_ind, nnz, num_row, row_ptr);                                                               \
        for (int64_t row = 0; row < num_row; row++) {                                                          \
            T *out = output + row * dim;                                                                       \
            for (int64_t d = 0; d < dim; d++) out[d] = nary_zero_##SUFFIX(sum);                                \
            for (int64_t ptr = row_ptr[row]; ptr < row_ptr[row + 1]; ptr++) {                                  \
                const T *rel = relation + edge_type[ptr] * dim;                                                \
                const T *in = input + col_ind[ptr] * dim;                                                      \
                T w = edge_weight[ptr];                                                                        \
                for (int64_t d = 0; d < dim; d++) {                                                            \
                    T x = bin_fwd_##SUFFIX(mul, rel[d], in[d]);                                                \
                    T y = w * x;                                                                               \
                    out[d] = nary_fwd_##SUFFIX(sum, out[d], y);                                                \
                }                                                                                              \
            }                                                                                                  \
        }                                                                                                      \
        free(row_ptr);                                                                                         \
        return 0;                                                                                              \
    }                                                                                                          \
    /* rspmm.cpp:77-119.  Gradients must be zero-initialised by the caller (zeros_like, rspmm.cpp:190-192). */ \
    int rspmm_oracle_backward_##SUFFIX(int sum, int mul, const int64_t *edge_index, const int64_t *edge_type,  \
                                       const T *edge_weight, const T *relation, const T *input,                \
                                       const T *output, const T *output_grad, T *weight_grad,                  \
                                       T *relation_grad, T *input_grad, int64_t num_row, int64_t nnz,          \
                                       int64_t dim) {                                                          \
        const int64_t *row_ind = edge_index, *col_ind = edge_index + nnz;                                      \
        int64_t *row_ptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(num_row + 1));                         \
        if (!row_ptr) return 1;                                                                                \
        ind2ptr(row_ind, nnz, num_row, row_ptr);                                                               \
        for (int64_t row = 0; row < num_row; row++) {                                                          \
            for (int64_t ptr = row_ptr[row]; ptr < row_ptr[row + 1]; ptr++) {                                  \
                int64_t col = col_ind[ptr], layer = edge_type[ptr];                                            \
                T w = edge_weight[ptr];                                                                        \
                T w_grad = 0;                                                                                  \
                for (int64_t d = 0; d < dim; d++) {                                                            \
                    T rel = relation[layer * dim + d];                                                         \
                    T in = input[col * dim + d];                                                               \
                    T out = output[row * dim + d];                                                             \
                    T out_grad = output_grad[row * dim + d];                                                   \
                    T x = bin_fwd_##SUFFIX(mul, rel, in);                                                      \
                    T y = w * x;                                                                               \
                    T dx_drel = bin_dlhs_##SUFFIX(mul, rel, in);                                               \
                    T dx_din = bin_drhs_##SUFFIX(mul, rel, in);                                                \
                    T dout_dy = nary_bwd_##SUFFIX(sum, out, y);                                                \
                    w_grad += out_grad * dout_dy * x;                                                          \
                    relation_grad[layer * dim + d] += out_grad * dout_dy * w * dx_drel;                        \
                    input_grad[col * dim + d] += out_grad * dout_dy * w * dx_din;                              \
                }                                                                                              \
                weight_grad[ptr] = w_grad;                                                                     \
            }                                                                                                  \
        }                                                                                                      \
        free(row_ptr);                                                                                         \
        return 0;                                                                                              \
    }

DEFINE_ORACLE(float, f32, FLT_MAX)
DEFINE_ORACLE(double, f64, DBL_MAX)
