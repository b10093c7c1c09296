/*
 * hll_matrix.h -- drop-in for reference libs/hll_matrix.h:12-39 (hacked ELLPACK, hack size 32).
 * Host layout is byte-identical to the reference: row-major slots r*MAXNZ + j inside each block,
 * last block short, JA = AS = NULL for an all-empty block.  Products run on the GPU (see
 * csr_matrix.h for the no-fallback rule).
 */
#ifndef SPMV_B200_HLL_MATRIX_H
#define SPMV_B200_HLL_MATRIX_H
#include <stddef.h>
#include "matrix_parser.h"
#ifdef __cplusplus
extern "C" {
#endif

#define HACK_SIZE 32

typedef struct {
    int M;      /* rows in this block            */
    int N;      /* columns of the matrix         */
    int MAXNZ;  /* longest row of the block      */
    int *JA;    /* [M*MAXNZ] column indices      */
    double *AS; /* [M*MAXNZ] coefficients        */
} ELLPACKBlock;

typedef struct {
    int num_blocks;
    ELLPACKBlock *blocks;
} HLLMatrix;

void init_hll_matrix(HLLMatrix *hll);
int convert_to_hll(const PreMatrix *pre, HLLMatrix *hll); /* reference src/hll_matrix.c:37-257 */
void free_hll_matrix(HLLMatrix *hll);
void printHLLMatrix(HLLMatrix *hll);
/* y[32 b + i] = sum_j AS*x[JA] (reference src/hll_matrix.c:286-308) -- GPU */
void spmv_hll_serial(int num_blocks, const ELLPACKBlock *blocks, const double *x, double *y);
/* reference src/hll_matrix.c:410-540 -- host */
int prepare_thread_distribution_hll(const HLLMatrix *matrix, int num_threads, int **thread_block_start,
                                    int **thread_block_end);
/* reference src/hll_matrix.c:339-408 -- GPU, only the blocks of the given ranges */
void spmv_hll(const ELLPACKBlock *blocks, const double *x, double *y, int num_threads,
              int const *thread_block_start, int const *thread_block_end);
void spmv_hll_simd(const ELLPACKBlock *blocks, const double *x, double *y, int num_threads,
                   int const *thread_block_start, int const *thread_block_end);

#ifdef __cplusplus
}
#endif
#endif
