/*
 * utility.c -- the helpers of reference src/utility.c that the CSR/HLL path uses:
 * init_vector_at_one (:18-22), the (col, value) row quicksort used by convert_in_csr (:25-91),
 * clear_cache (:141-159) and process_matrix_file (:160-172).
 */
#include "utility.h"

#include <stdio.h>
#include <string.h>

void init_vector_at_one(double *v, const int size) {
    for (int i = 0; i < size; ++i) v[i] = 1.0;
}

void swap(int *a, int *b) {
    int t = *a;
    *a = *b;
    *b = t;
}

void swap_double(double *a, double *b) {
    double t = *a;
    *a = *b;
    *b = t;
}

/* Lomuto scheme, pivot = last element, "<=" predicate.  convert_in_csr keeps duplicate
 * (row, col) entries, and the order in which their VALUES end up is decided by exactly this
 * swap sequence, so it is part of the bit-exact CSR contract. */
size_t partition(int *col_idx, double *values, size_t low, size_t high) {
    const int pivot = col_idx[high];
    size_t store = low;
    for (size_t k = low; k < high; ++k) {
        if (col_idx[k] <= pivot) {
            swap(&col_idx[store], &col_idx[k]);
            swap_double(&values[store], &values[k]);
            ++store;
        }
    }
    swap(&col_idx[store], &col_idx[high]);
    swap_double(&values[store], &values[high]);
    return store;
}

/* Sorts [low, high] (both inclusive).  Sub-ranges are disjoint, so the visiting order does not
 * change the result; a growable explicit stack replaces the reference's recursion / fixed
 * 64-slot stack (which can overflow on adversarial rows longer than 10 000). */
void sort_row(int *col_idx, double *values, size_t low, size_t high) {
    if (low >= high) return;
    size_t cap = 64, top = 0;
    size_t *stack = malloc(cap * sizeof(size_t));
    if (!stack) return;
    stack[top++] = low;
    stack[top++] = high;
    while (top) {
        const size_t hi = stack[--top], lo = stack[--top];
        const size_t p = partition(col_idx, values, lo, hi);
        if (top + 4 > cap) {
            size_t *grown = realloc(stack, 2 * cap * sizeof(size_t));
            if (!grown) break;
            stack = grown;
            cap *= 2;
        }
        if (p > lo + 1) { stack[top++] = lo; stack[top++] = p - 1; }
        if (p + 1 < hi) { stack[top++] = p + 1; stack[top++] = hi; }
    }
    free(stack);
}

void clear_cache(size_t clear_size_mb) {
    const size_t bytes = clear_size_mb << 20;
    volatile char *scratch = malloc(bytes ? bytes : 1);
    if (!scratch) return;
    for (size_t i = 0; i < bytes; i += 64) scratch[i] = (char)i;
    free((void *)scratch);
}

int process_matrix_file(const char *filepath, PreMatrix *pre_mat) {
    init_pre_matrix(pre_mat);
    if (read_matrix_market(filepath, pre_mat) != 0) {
        printf("process_matrix_file: could not read '%s'\n", filepath);
        return -1;
    }
    return 0;
}
