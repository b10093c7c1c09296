/*
 * oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the reference's CSR / HLL y = A*x path
 * (MarcoLor01/SparseMatrixVectorMultiplication).  Every function cites the reference
 * file:line whose behaviour it restates.  Parity is PINNED: tests/test_oracle_pinned.py checks
 * this restatement against (a) the golden vectors in tests/golden/ that were produced by the
 * unmodified reference sources (tests/golden/make_golden.py) and (b) oracle/_ref/libspmv_ref.so
 * -- the unmodified reference compiled by oracle/Makefile -- on randomised inputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (sparsematrixvectormultiplication_b200/) never does.
 */
#ifndef SPMV_ORACLE_H
#define SPMV_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_HACK_SIZE 32 /* reference libs/hll_matrix.h:12 */

/* COO container, restating PreMatrix (reference libs/matrix_parser.h:6-14). */
typedef struct {
    int M, N, nz;
    int *I, *J;
    double *val;
    char type[4]; /* MM_typecode: reference libs/mmio.h:21 */
} orc_coo;

/* HLL restated as one flat row-major arena (the reference mallocs JA/AS per block,
 * libs/hll_matrix.h:15-27; contents and in-block order are identical). */
typedef struct {
    int num_blocks;
    int *rows;          /* [num_blocks]   ELLPACKBlock.M  (last block short)        */
    int *maxnz;         /* [num_blocks]   ELLPACKBlock.MAXNZ                         */
    long long *offset;  /* [num_blocks+1] start of block b inside JA / AS            */
    int *JA;            /* row-major inside a block: JA[offset[b] + r*maxnz[b] + j]  */
    double *AS;
} orc_hll;

/* reference src/matrix_parser.c:25-150 (+ libs/mmio.c:96-214). 0 ok, -1 error. */
int orc_read_matrix_market(const char *path, orc_coo *out);
void orc_free_coo(orc_coo *c);

/* reference src/csr_matrix.c:63-126 + src/utility.c:38-91. Caller provides
 * row_ptr[M+1], col_idx[nz], values[nz]. */
int orc_coo_to_csr(int M, int nz, const int *I, const int *J, const double *val,
                   int *row_ptr, int *col_idx, double *values);

/* reference src/hll_matrix.c:37-257. */
int orc_coo_to_hll(int M, int N, int nz, const int *I, const int *J, const double *val,
                   orc_hll *out);
void orc_free_hll(orc_hll *h);

/* reference src/csr_matrix.c:130-139 : y[i] += sum (caller zeroes y). */
void orc_spmv_csr_serial(int M, const int *row_ptr, const int *col_idx, const double *values,
                         const double *x, double *y);
/* reference src/hll_matrix.c:286-308 : y[32*b+i] = sum, padding slots included. */
void orc_spmv_hll_serial(const orc_hll *h, const double *x, double *y);

/* reference src/csr_matrix.c:167-266 (greedy nnz-balanced contiguous row ranges).
 * start/end must hold num_threads ints. Returns the number of ranges used. */
int orc_partition_rows(int M, const int *row_ptr, int num_threads, long long total_nnz,
                       int *start, int *end);
/* reference src/hll_matrix.c:410-540 (same greedy rule over blocks, weight = slots whose
 * JA is a valid column, i.e. rows*MAXNZ). */
int orc_partition_hll_blocks(const orc_hll *h, int N, int num_threads, int *start, int *end);

/* reference src/csr_matrix.c:294-313 (y[i] = sum per OpenMP thread range). */
void orc_spmv_csr_parallel(const int *row_ptr, const int *col_idx, const double *values,
                           const double *x, double *y, int num_threads,
                           const int *start, const int *end);
/* reference src/hll_matrix.c:376-408. */
void orc_spmv_hll_parallel(const orc_hll *h, const double *x, double *y, int num_threads,
                           const int *start, const int *end);

/* reference src/performance_calculate.c:98-101 and :116-178 (C build) /
 * cuda_src/performance_calculate.cu:103-148 (CUDA build). */
double orc_calculate_flops(int nz, double seconds);
int orc_diff_metrics_c(const double *ref, const double *res, int n, double abs_tol,
                       double rel_tol, double *mean_rel_err);
void orc_diff_metrics_cuda(const double *ref, const double *res, int n,
                           double *mean_abs_err, double *mean_rel_err);

/* BASELINE.json config 5 semantics (no reference code: the reference only repeats the same
 * product): repeat iters times  y = A x ; lambda = ||y||_2 ;
 * x = y / lambda.  The product is the serial CSR loop above.  x is updated in place,
 * lambdas[iters] receives every norm. */
void orc_power_iteration(int M, const int *row_ptr, const int *col_idx, const double *values,
                         double *x, double *y, int iters, double *lambdas);

/* bench.py --impl reference at N > 1: the full-size 7-point Laplacian written by all host threads (returns nnz; with
 * col_idx == NULL only row_ptr is filled) and the norm + scale of one power iteration in an OpenMP region. */
long long orc_lap3d_csr(int n, int *row_ptr, int *col_idx, double *values);
double orc_norm_scale(const double *y, double *x, long long n, int threads);

#ifdef __cplusplus
}
#endif
#endif
