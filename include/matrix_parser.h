/*
 * matrix_parser.h -- drop-in for reference libs/matrix_parser.h:6-19 (same struct layout,
 * same signatures, same 0 / -1 error convention with a message on stdout).
 */
#ifndef SPMV_B200_MATRIX_PARSER_H
#define SPMV_B200_MATRIX_PARSER_H
#include <stdbool.h>
#include "mmio.h"
#ifdef __cplusplus
extern "C" {
#endif

/* COO as read from a Matrix Market file: 0-based, symmetric entries mirrored, pattern -> 1.0 */
typedef struct {
    int M;
    int N;
    int nz;
    int *I;
    int *J;
    double *val;
    MM_typecode type;
} PreMatrix;

void init_pre_matrix(PreMatrix *mat);
void free_pre_matrix(PreMatrix *mat);
int read_matrix_market(const char *filename, PreMatrix *mat);
void print_pre_matrix(PreMatrix *mat, const bool full_print);

#ifdef __cplusplus
}
#endif
#endif
