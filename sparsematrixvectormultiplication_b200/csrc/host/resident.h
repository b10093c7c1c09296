/* resident.h -- opt-in cache of device copies behind the stateless drop-in product signatures (see spmv_b200.h). */
#ifndef SPMV_B200_RESIDENT_H
#define SPMV_B200_RESIDENT_H
#include "hll_matrix.h"
#include "spmv_b200.h"

/* a resident handle for these host arrays: the cached one when it still matches, else a fresh upload (which replaces
 * the cached one when the cache is on).  *borrowed = 1: the cache owns the handle, do not free it. */
int resident_csr(int M, int N, long long nnz, const int *row_ptr, const int *col_idx, const double *values,
                 spmv_b200_csr **out, int *borrowed);
int resident_hll(const ELLPACKBlock *blocks, int count, int rows, int N, spmv_b200_hll **out, int *borrowed);
void resident_forget_csr(const void *any_array);   /* free_csr_matrix: drop if the cached copy came from these arrays */
void resident_forget_hll(const void *blocks);
#endif
