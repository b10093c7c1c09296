/*
 * csr_matrix.h -- drop-in for reference libs/csr_matrix.h:8-33.
 *
 * Builders run on the host and are bit-exact with the reference.  The three product entry
 * points keep their signatures but execute on the GPU through the C-ABI in spmv_b200.h
 * (upload, sm_100a kernel, download).  There is no CPU fallback: if no CUDA device is usable
 * they print the error on stderr, fill the rows they own in y with NaN and leave the reason
 * in spmv_b200_last_error().
 */
#ifndef SPMV_B200_CSR_MATRIX_H
#define SPMV_B200_CSR_MATRIX_H
#include <stddef.h>
#include "matrix_parser.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int M;
    int N;
    int nz;
    int *row_ptr;
    int *col_idx;
    double *values;
    MM_typecode type;
} CSRMatrix;

void init_csr_matrix(CSRMatrix *mat);
void free_csr_matrix(CSRMatrix *mat);
/* reference src/csr_matrix.c:63-126 (matrix_name is unused, as in the reference) */
int convert_in_csr(const PreMatrix *pre, CSRMatrix *csr, const char *matrix_name);
void print_csr_matrix(const CSRMatrix *mat);
/* reference libs/csr_matrix.h:23, src/csr_matrix.c:28-61 */
void write_memory_stats_to_csv(const char *matrix_name, int nz, size_t total_memory_bytes);

/* y += A x (reference src/csr_matrix.c:130-139) -- GPU */
void csr_matrix_vector_mult(int num_row, const int *row_ptr, const int *col_idx, const double *values,
                            const double *x, double *y);
/* reference src/csr_matrix.c:167-266 -- host; also the multi-GPU row partitioner */
int prepare_thread_distribution(const int num_row, const int *row_ptr, int num_threads,
                                const long long total_nnz, int **thread_row_start, int **thread_row_end);
/* y[i] = (A x)[i] for the rows of every range (reference src/csr_matrix.c:269-313) -- GPU */
void spvm_csr_parallel(const int *row_ptr, const int *col_idx, const double *values, const double *x,
                       double *y, int num_threads, const int *thread_row_start, const int *thread_row_end);
void spvm_csr_parallel_simd(const int *row_ptr, const int *col_idx, const double *values, const double *x,
                            double *y, int num_threads, const int *thread_row_start,
                            const int *thread_row_end);

#ifdef __cplusplus
}
#endif
#endif
