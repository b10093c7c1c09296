/*
 * utility.c -- the helpers of reference src/utility.c that the CSR/HLL path and the reference's
 * own OpenMP driver (main.c) link against: init_vector_at_one (:18-22), the (col, value) row
 * quicksort used by convert_in_csr (:25-91), write_results_to_csv (:95-138), clear_cache
 * (:141-159), process_matrix_file (:160-172) and create_directory (:200-216).
 */
#define _DEFAULT_SOURCE /* lstat */
#include "utility.h"

#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

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
 * change the result: the smaller side is handled by recursion, the larger one by the loop, which
 * bounds the depth by log2(n) without any allocation (the reference recurses on both sides for
 * spans up to 10 000 and uses a fixed 64-slot stack above that, which adversarial rows can overflow). */
void sort_row(int *col_idx, double *values, size_t low, size_t high) {
    while (low < high) {
        const size_t p = partition(col_idx, values, low, high);
        const size_t left = p - low, right = high - p;
        if (left < right) {
            if (p > low + 1) sort_row(col_idx, values, low, p - 1);
            low = p + 1;
        } else {
            if (p + 1 < high) sort_row(col_idx, values, p + 1, high);
            if (p == 0) break;
            high = p - 1;
        }
    }
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

/* One row of the OpenMP driver's result file (reference src/utility.c:95-138, called from main.c:441-450).  The file
 * is opened for appending; an empty file first receives the header.  Column names, column order and the number
 * format ("%.15f") are the reference's -- they are the contract of whatever reads the result CSVs afterwards. */
void write_results_to_csv(const char *matrix_name, const int num_rows, const int num_cols, const int nz,
                          const int num_threads, const double time_serial, const double time_serial_hll,
                          const double time_parallel, const double time_parallel_simd, const double time_parallel_hll,
                          const double time_parallel_hll_simd, DiffMetrics error_csr, DiffMetrics error_hll,
                          DiffMetrics error_csr_simd, DiffMetrics error_hll_simd, const double speedup_parallel,
                          const double speedup_simd, const double speedup_hll, const double speedup_hll_simd,
                          const double efficiency_parallel, const double efficiency_simd, const double efficiency_hll,
                          const double efficiency_hll_simd, const double flops_serial, const double avg_flops_hll_serial,
                          const double flops_parallel, const double flops_parallel_simd, const double flops_parallel_hll,
                          const double flops_parallel_hll_simd, const char *output_file) {
    const struct {
        const char *name;
        double value;
    } columns[] = {
        {"time_serial", time_serial},
        {"time_serial_hll", time_serial_hll},
        {"time_parallel", time_parallel},
        {"time_parallel_simd", time_parallel_simd},
        {"time_parallel_hll", time_parallel_hll},
        {"time_parallel_hll_simd", time_parallel_hll_simd},
        {"error_csr_relative", error_csr.mean_rel_err},
        {"error_csr_absolute", error_csr.mean_abs_err},
        {"error_hll_relative", error_hll.mean_rel_err},
        {"error_hll_absolute", error_hll.mean_abs_err},
        {"error_csr_simd_relative", error_csr_simd.mean_rel_err},
        {"error_csr_simd_absolute", error_csr_simd.mean_abs_err},
        {"error_hll_simd_relative", error_hll_simd.mean_rel_err},
        {"error_hll_simd_absolute", error_hll_simd.mean_abs_err},
        {"flops_serial", flops_serial},
        {"flops_serial_hll", avg_flops_hll_serial},
        {"flops_parallel", flops_parallel},
        {"flops_parallel_simd", flops_parallel_simd},
        {"flops_parallel_hll", flops_parallel_hll},
        {"flops_parallel_hll_simd", flops_parallel_hll_simd},
        {"speedup_parallel", speedup_parallel},
        {"speedup_simd", speedup_simd},
        {"speedup_hll", speedup_hll},
        {"speedup_hll_simd", speedup_hll_simd},
        {"efficiency_parallel", efficiency_parallel},
        {"efficiency_simd", efficiency_simd},
        {"efficiency_hll", efficiency_hll},
        {"efficiency_hll_simd", efficiency_hll_simd},
    };
    const size_t ncols = sizeof columns / sizeof columns[0];
    FILE *out = fopen(output_file, "a+");
    if (!out) {
        printf("write_results_to_csv: cannot open %s (%s)\n", output_file, strerror(errno));
        return;
    }
    fseek(out, 0, SEEK_END);
    if (ftell(out) == 0) {
        fputs("matrix_name,rows,cols,nonzeros,num_threads", out);
        for (size_t c = 0; c < ncols; ++c) fprintf(out, ",%s", columns[c].name);
        fputc('\n', out);
    }
    fprintf(out, "%s,%d,%d,%d,%d", matrix_name, num_rows, num_cols, nz, num_threads);
    for (size_t c = 0; c < ncols; ++c) fprintf(out, ",%.15f", columns[c].value);
    fputc('\n', out);
    fclose(out);
}

/* The reference's create_directory (src/utility.c:174-216) makes `path` or, if it already exists, DELETES every
 * entry in it, and exit()s on failure.  Here: the directory is created when missing and otherwise left alone -- a
 * library must not wipe a caller's directory as a side effect, and write_results_to_csv appends anyway.  Setting
 * SPMV_B200_WIPE_RESULT_DIR=1 restores the reference's behaviour for regular files (sub-directories are never
 * touched).  Failures are reported on stdout and do not end the process. */
void create_directory(const char *path) {
    if (!path || !*path) return;
    if (mkdir(path, 0777) == 0) {
        printf("create_directory: created '%s'\n", path);
        return;
    }
    if (errno != EEXIST) {
        printf("create_directory: cannot create '%s' (%s)\n", path, strerror(errno));
        return;
    }
    const char *wipe = getenv("SPMV_B200_WIPE_RESULT_DIR");
    if (!wipe || strcmp(wipe, "1") != 0) {
        printf("create_directory: '%s' exists; its content is kept (SPMV_B200_WIPE_RESULT_DIR=1 empties it)\n", path);
        return;
    }
    DIR *dir = opendir(path);
    if (!dir) {
        printf("create_directory: cannot list '%s' (%s)\n", path, strerror(errno));
        return;
    }
    for (struct dirent *e = readdir(dir); e; e = readdir(dir)) {
        char file[4096];
        struct stat st;
        if (snprintf(file, sizeof file, "%s/%s", path, e->d_name) >= (int)sizeof file) continue;
        if (lstat(file, &st) == 0 && S_ISREG(st.st_mode) && unlink(file) != 0)
            printf("create_directory: cannot remove '%s' (%s)\n", file, strerror(errno));
    }
    closedir(dir);
}
