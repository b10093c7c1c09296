/*
 * hll_matrix.c -- drop-in for reference src/hll_matrix.c.
 *
 *   convert_to_hll                   host, byte-identical blocks to reference :37-257
 *   prepare_thread_distribution_hll  host, same block ranges as reference :410-540
 *   spmv_hll_serial, spmv_hll[_simd] same signatures as reference :286-308 / :339-408, executed on
 *                                    the GPU through spmv_b200.h.  No CPU fallback.
 *
 * The builder differs from the reference in mechanics only: one counting pass lays all entries
 * out row by row in a single arena (instead of M small mallocs), rows are ordered by a stable
 * insertion / merge sort (glibc's qsort is a stable merge sort, so duplicates keep file order)
 * and blocks are filled in parallel.
 */
#include "hll_matrix.h"

#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "resident.h"
#include "spmv_b200.h"
#include "utility.h"

void init_hll_matrix(HLLMatrix *hll) {
    hll->num_blocks = 0;
    hll->blocks = NULL;
}

void free_hll_matrix(HLLMatrix *hll) {
    if (!hll) return;
    resident_forget_hll(hll->blocks); /* a cached device copy of these blocks dies with them */
    if (hll->blocks) {
        for (int b = 0; b < hll->num_blocks; ++b) {
            FREE_CHECK(hll->blocks[b].JA);
            FREE_CHECK(hll->blocks[b].AS);
        }
        FREE_CHECK(hll->blocks);
    }
    hll->num_blocks = 0;
}

typedef struct {
    int col;
    double val;
} Entry;

/* stable: equal columns keep their arrival (file) order */
static void stable_order(Entry *e, Entry *tmp, int n) {
    if (n <= 24) {
        for (int k = 1; k < n; ++k) {
            const Entry held = e[k];
            int p = k;
            while (p > 0 && e[p - 1].col > held.col) {
                e[p] = e[p - 1];
                --p;
            }
            e[p] = held;
        }
        return;
    }
    const int half = n / 2;
    stable_order(e, tmp, half);
    stable_order(e + half, tmp, n - half);
    if (e[half - 1].col <= e[half].col) return;
    int a = 0, b = half, o = 0;
    while (a < half && b < n) tmp[o++] = e[b].col < e[a].col ? e[b++] : e[a++];
    while (a < half) tmp[o++] = e[a++];
    while (b < n) tmp[o++] = e[b++];
    memcpy(e, tmp, (size_t)n * sizeof(Entry));
}

int convert_to_hll(const PreMatrix *pre, HLLMatrix *hll) {
    if (!hll) {
        printf("convert_to_hll: invalid arguments\n");
        return -1;
    }
    init_hll_matrix(hll);
    if (!pre) {
        printf("convert_to_hll: invalid arguments\n");
        return -1;
    }
    const int M = pre->M, N = pre->N, nz = pre->nz;
    const int nb = (M + HACK_SIZE - 1) / HACK_SIZE;
    hll->num_blocks = nb;
    hll->blocks = calloc((size_t)(nb > 0 ? nb : 1), sizeof(ELLPACKBlock));
    long long *first = calloc((size_t)M + 1, sizeof(long long)); /* arena offset of each row */
    int *fill = calloc((size_t)(M > 0 ? M : 1), sizeof(int));
    Entry *arena = malloc((size_t)(nz > 0 ? nz : 1) * sizeof(Entry));
    Entry *scratch = malloc((size_t)(nz > 0 ? nz : 1) * sizeof(Entry));
    int status = (hll->blocks && first && fill && arena && scratch) ? 0 : -1;
    if (status != 0) printf("convert_to_hll: out of memory\n");

    /* Counting scatter by row.  Large, well-formed inputs: every thread owns a range of rows and scans the whole COO
     * list for them (sequential reads; the entries of a row keep their file order, so the result is bit-identical to
     * the serial loops below, which remain for small inputs and for anything with an out-of-range index, whose
     * messages must come out in file order). */
    int threads = 1;
#ifdef _OPENMP
    if (status == 0 && nz >= (1 << 20) && M >= 64) {
        int clean = 1;
#pragma omp parallel for reduction(& : clean)
        for (int k = 0; k < nz; ++k) clean &= pre->I[k] >= 0 && pre->I[k] < M && pre->J[k] >= 0 && pre->J[k] < N;
        if (clean) threads = omp_get_max_threads() < 64 ? omp_get_max_threads() : 64;
    }
#endif
    if (threads > 1) {
#pragma omp parallel num_threads(threads)
        {
            const int t = omp_get_thread_num(), T = omp_get_num_threads();
            const int r0 = (int)((long long)M * t / T), r1 = (int)((long long)M * (t + 1) / T);
            for (int k = 0; k < nz; ++k) {
                const int r = pre->I[k];
                if (r >= r0 && r < r1) first[r + 1]++;
            }
        }
        for (int r = 0; r < M; ++r) first[r + 1] += first[r];
#pragma omp parallel num_threads(threads)
        {
            const int t = omp_get_thread_num(), T = omp_get_num_threads();
            const long long lo_target = (long long)nz * t / T, hi_target = (long long)nz * (t + 1) / T;
            int a = 0, b = M;
            while (a < b) { const int m = (a + b) / 2; if (first[m] < lo_target) a = m + 1; else b = m; }
            const int r0 = t == 0 ? 0 : a;
            a = 0; b = M;
            while (a < b) { const int m = (a + b) / 2; if (first[m] < hi_target) a = m + 1; else b = m; }
            const int r1 = t == T - 1 ? M : a;
            for (int k = 0; k < nz; ++k) {
                const int r = pre->I[k];
                if (r >= r0 && r < r1) {
                    Entry *slot = &arena[first[r] + fill[r]++];
                    slot->col = pre->J[k];
                    slot->val = pre->val[k];
                }
            }
        }
    } else {
    for (int k = 0; status == 0 && k < nz; ++k) {
        const int r = pre->I[k];
        if (r < 0 || r >= M) {
            printf("convert_to_hll: row index %d outside [0,%d)\n", r, M);
            status = -1;
        } else {
            first[r + 1]++;
        }
    }
    if (status == 0) {
        for (int r = 0; r < M; ++r) first[r + 1] += first[r];
        for (int k = 0; k < nz; ++k) {
            const int r = pre->I[k], c = pre->J[k];
            if (c < 0 || c >= N) { /* the reference warns and skips such entries (:193-196) */
                printf("convert_to_hll: skipping entry with column %d outside [0,%d)\n", c, N);
                continue;
            }
            Entry *slot = &arena[first[r] + fill[r]++];
            slot->col = c;
            slot->val = pre->val[k];
        }
    }
    }

    int failed = 0;
    if (status == 0) {
#pragma omp parallel for schedule(dynamic, 64) reduction(| : failed)
        for (int b = 0; b < nb; ++b) {
            const int r0 = b * HACK_SIZE;
            const int rows = (b == nb - 1) ? M - r0 : HACK_SIZE; /* the last block is short */
            int width = 0;
            for (int r = r0; r < r0 + rows; ++r) {
                const int len = (int)(first[r + 1] - first[r]); /* counted, incl. skipped ones */
                if (len > width) width = len;
            }
            ELLPACKBlock *blk = &hll->blocks[b];
            blk->M = rows;
            blk->N = N;
            blk->MAXNZ = width;
            blk->JA = NULL;
            blk->AS = NULL;
            if (width == 0) continue; /* empty block: NULL arrays, as the reference */
            blk->JA = malloc((size_t)width * rows * sizeof(int));
            blk->AS = malloc((size_t)width * rows * sizeof(double));
            if (!blk->JA || !blk->AS) {
                failed |= 1;
                continue;
            }
            for (int lr = 0; lr < rows; ++lr) {
                const int r = r0 + lr;
                Entry *e = arena + first[r];
                const int len = fill[r];
                stable_order(e, scratch + first[r], len);
                int *ja = blk->JA + (size_t)lr * width;
                double *as = blk->AS + (size_t)lr * width;
                int pad_col = 0; /* empty rows are padded with column 0 */
                for (int j = 0; j < len; ++j) {
                    ja[j] = pad_col = e[j].col;
                    as[j] = e[j].val;
                }
                for (int j = len; j < width; ++j) { /* padding repeats the last real column */
                    ja[j] = pad_col;
                    as[j] = 0.0;
                }
            }
        }
    }
    if (failed) {
        printf("convert_to_hll: out of memory while filling blocks\n");
        status = -1;
    }
    free(first);
    free(fill);
    free(arena);
    free(scratch);
    if (status != 0) free_hll_matrix(hll);
    return status;
}

void printHLLMatrix(HLLMatrix *hll) {
    printf("HLL matrix with %d blocks\n", hll->num_blocks);
    for (int b = 0; b < hll->num_blocks; ++b) {
        const ELLPACKBlock *blk = &hll->blocks[b];
        printf("block %d: %d rows, %d columns, MAXNZ=%d\n", b, blk->M, blk->N, blk->MAXNZ);
        for (int r = 0; r < blk->M; ++r) {
            printf("  row %d:", r);
            for (int j = 0; j < blk->MAXNZ; ++j)
                printf(" (%d, %.6f)", blk->JA[r * blk->MAXNZ + j], blk->AS[r * blk->MAXNZ + j]);
            printf("\n");
        }
    }
}

/* Same greedy rule as prepare_thread_distribution, over blocks; a block weighs the number of
 * its slots whose JA is a valid column (padding repeats a valid column, so this is
 * rows*MAXNZ on well-formed data -- reference :444-462). */
int prepare_thread_distribution_hll(const HLLMatrix *matrix, int num_threads, int **thread_block_start,
                                    int **thread_block_end) {
    if (!matrix || num_threads <= 0 || !thread_block_start || !thread_block_end) {
        printf("prepare_thread_distribution_hll: invalid arguments\n");
        return 0;
    }
    const int nb = matrix->num_blocks;
    if (num_threads > nb) num_threads = nb;
    int *start = malloc((size_t)(num_threads > 0 ? num_threads : 1) * sizeof(int));
    int *end = malloc((size_t)(num_threads > 0 ? num_threads : 1) * sizeof(int));
    long long *weight = malloc((size_t)(nb > 0 ? nb : 1) * sizeof(long long));
    if (!start || !end || !weight) {
        free(start);
        free(end);
        free(weight);
        printf("prepare_thread_distribution_hll: out of memory\n");
        *thread_block_start = *thread_block_end = NULL;
        return 0;
    }
    long long total = 0;
#pragma omp parallel for reduction(+ : total)
    for (int b = 0; b < nb; ++b) {
        const ELLPACKBlock *blk = &matrix->blocks[b];
        long long valid = 0;
        const long long slots = (long long)blk->M * blk->MAXNZ;
        if (blk->JA)
            for (long long s = 0; s < slots; ++s) valid += (blk->JA[s] >= 0 && blk->JA[s] < blk->N);
        weight[b] = valid;
        total += valid;
    }
    int used = 0;
    if (num_threads > 0) {
        const long long target = (total + num_threads - 1) / num_threads;
        int part = 0, first = 0;
        long long gathered = 0;
        for (int b = 0; b < nb; ++b) {
            gathered += weight[b];
            const int close = (part != num_threads - 1 && gathered >= target) || b == nb - 1;
            if (!close) continue;
            if (gathered > 0) {
                start[used] = first;
                end[used] = b + 1;
                ++used;
            }
            first = b + 1;
            gathered = 0;
            ++part;
        }
    }
    free(weight);
    *thread_block_start = start;
    *thread_block_end = end;
    return used;
}

/* ------------------------------------------------------------------------------------------
 * GPU-backed product entry points.
 * ---------------------------------------------------------------------------------------- */
static int run_blocks(const ELLPACKBlock *blocks, int count, const double *x, double *y_out) {
    if (count <= 0) return SPMV_B200_OK;
    long long rows = 0;
    for (int b = 0; b < count; ++b) rows += blocks[b].M;
    spmv_b200_hll *H = NULL;
    int borrowed = 0;
    int rc = resident_hll(blocks, count, (int)rows, blocks[0].N, &H, &borrowed);
    if (rc == SPMV_B200_OK) rc = spmv_b200_hll_spmv_host(H, x, y_out);
    if (!borrowed) spmv_b200_hll_free(H);
    if (rc != SPMV_B200_OK) {
        fprintf(stderr, "spmv_hll: GPU product failed: %s\n", spmv_b200_last_error());
        for (long long i = 0; i < rows; ++i) y_out[i] = NAN;
    }
    return rc;
}

void spmv_hll_serial(int num_blocks, const ELLPACKBlock *blocks, const double *x, double *y) {
    run_blocks(blocks, num_blocks, x, y);
}

/* One upload and one product for the union of the per-thread block ranges (the reference's partitions tile
 * [0, num_blocks) without gaps, src/hll_matrix.c:471-498); only rows inside a range are written. */
static void ranged_blocks(const ELLPACKBlock *blocks, const double *x, double *y, int num_threads,
                          const int *lo, const int *hi) {
    if (num_threads <= 0) return;
    int first = lo[0], last = hi[0];
    for (int t = 1; t < num_threads; ++t) {
        if (lo[t] < first) first = lo[t];
        if (hi[t] > last) last = hi[t];
    }
    if (last <= first) return;
    if (first != 0) { /* unusual: ranges that do not start at block 0 keep the simple path */
        for (int t = 0; t < num_threads; ++t) run_blocks(blocks + lo[t], hi[t] - lo[t], x, y + (size_t)lo[t] * HACK_SIZE);
        return;
    }
    double *full = malloc((size_t)last * HACK_SIZE * sizeof(double));
    if (!full) {
        for (int t = 0; t < num_threads; ++t)
            for (size_t i = (size_t)lo[t] * HACK_SIZE; i < (size_t)hi[t] * HACK_SIZE; ++i) y[i] = NAN;
        fprintf(stderr, "spmv_hll: out of host memory\n");
        return;
    }
    run_blocks(blocks, last, x, full);
    for (int t = 0; t < num_threads; ++t) {
        size_t rows = 0;
        for (int b = lo[t]; b < hi[t]; ++b) rows += (size_t)blocks[b].M;
        memcpy(y + (size_t)lo[t] * HACK_SIZE, full + (size_t)lo[t] * HACK_SIZE, rows * sizeof(double));
    }
    free(full);
}

void spmv_hll(const ELLPACKBlock *blocks, const double *x, double *y, int num_threads,
              int const *thread_block_start, int const *thread_block_end) {
    ranged_blocks(blocks, x, y, num_threads, thread_block_start, thread_block_end);
}

void spmv_hll_simd(const ELLPACKBlock *blocks, const double *x, double *y, int num_threads,
                   int const *thread_block_start, int const *thread_block_end) {
    ranged_blocks(blocks, x, y, num_threads, thread_block_start, thread_block_end);
}
