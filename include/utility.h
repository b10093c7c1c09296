/*
 * utility.h -- the part of reference libs/utility.h:7-31 that the CSR/HLL path needs:
 * ITERATION_SKIP, FREE_CHECK, x = 1 initialisation, the row quicksort used by convert_in_csr
 * and the file loader.  The CSV writers and the directory wipe of the reference's drivers are
 * out of scope.
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
