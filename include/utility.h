/*
 * utility.h -- the part of reference libs/utility.h:7-31 that the CSR/HLL path needs:
 * ITERATION_SKIP, FREE_CHECK, x = 1 initialisation, the row quicksort used by convert_in_csr,
 * the file loader, and the two helpers the reference's OpenMP driver calls (main.c:29,441):
 * write_results_to_csv and create_directory -- so that main.c relinks against this library
 * unchanged (tests/test_relink_reference_main.py).  create_directory does NOT wipe an existing
 * directory unless SPMV_B200_WIPE_RESULT_DIR=1 (the reference always does, src/utility.c:200-216).
 */
#ifndef SPMV_B200_UTILITY_H
#define SPMV_B200_UTILITY_H
#include <stddef.h>
#include <stdlib.h>
#include "matrix_parser.h"
#include "performance_calculate.h"
#ifdef __cplusplus
extern "C" {
#endif

#define ITERATION_SKIP 5
#define FREE_CHECK(ptr)        \
    do {                       \
        if ((ptr) != NULL) {   \
            free(ptr);         \
            (ptr) = NULL;      \
        }                      \
    } while (0)

void init_vector_at_one(double *v, const int size);
/* reference libs/utility.h:17-23, src/utility.c:95-138 */
void write_results_to_csv(const char *matrix_name, const int num_rows, const int num_cols, const int nz,
                          const int num_threads, const double time_serial, const double time_serial_hll, const double time_parallel,
                          const double time_parallel_simd, const double time_parallel_hll, const double time_parallel_hll_simd,
                          DiffMetrics error_csr, DiffMetrics error_hll, DiffMetrics error_csr_simd, DiffMetrics error_hll_simd, const double speedup_parallel,
                          const double speedup_simd, const double speedup_hll, const double speedup_hll_simd, const double efficiency_parallel,
                          const double efficiency_simd, const double efficiency_hll, const double efficiency_hll_simd, const double flops_serial, const double avg_flops_hll_serial,
                          const double flops_parallel, const double flops_parallel_simd, const double flops_parallel_hll, const double flops_parallel_hll_simd, const char *output_file);
/* reference libs/utility.h:29, src/utility.c:200-216 (see the note above: no wipe by default, no exit()) */
void create_directory(const char *path);
void swap(int *a, int *b);
void swap_double(double *a, double *b);
size_t partition(int *col_idx, double *values, size_t low, size_t high);
void sort_row(int *col_idx, double *values, size_t low, size_t high);
void clear_cache(size_t clear_size_mb);
int process_matrix_file(const char *filepath, PreMatrix *pre_mat);

#ifdef __cplusplus
}
#endif
#endif
