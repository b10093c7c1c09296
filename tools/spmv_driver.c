/*
 * spmv_driver.c -- a main()-style driver for the B200 engine, shaped like the reference's CUDA driver
 * (reference main_cuda.cu:40-720) but written against the drop-in host API and the C-ABI only:
 *
 *   for every .mtx file:  read_matrix_market -> convert_in_csr -> convert_to_hll            (host, reference-exact)
 *                         spmv_b200_csr_upload / spmv_b200_hll_upload                        (once, as main_cuda.cu:128-145)
 *                         ITERATION_SKIP warm-up + 95 timed products per kernel, cudaEvents   (main_cuda.cu:159-200)
 *                         every result checked against the serial-order CSR product          (computeDifferenceMetrics)
 *                         one CSV row                                                         (cuda_src/utility.cu:93-133)
 *
 * CSV columns: the reference's leading columns (matrix_name, rows, cols, nonzeros) followed, per kernel, by
 * time_* (mean seconds), flops_* (2 nz / t), gbs_* (algorithmic bytes / t), roofline_* (fraction of --peak-gbs)
 * and relative_/absolute_error_*; then the host-buffer (end-to-end) times.
 *
 * Usage: spmv_driver [--csv out.csv] [--iters 95] [--warmup 5] [--peak-gbs 8000] [--gpus N] [--lap2d n] matrix.mtx ...
 *        --lap2d n  adds the 2-D 5-point Laplacian on an n x n grid, generated on the device (no file).
 *        --gpus N   (N >= 1; square matrices from files) additionally runs the row-partitioned product and 20 power
 *                   iterations on N GPUs through spmv_b200_multi_* (fused MAILBOX exchange when every row is short enough,
 *                   else ALLGATHER_PEER): columns ngpus, time_multi_iteration, flops_multi_iteration, lambda_multi,
 *                   relative_error_multi_product.  Without it those columns read 1, 0, 0, 0, 0.
 * Exit status: 0 ok, 1 usage / file error, 2 a result differs from the serial-order product, 3 no CUDA device.
 */
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "csr_matrix.h"
#include "hll_matrix.h"
#include "matrix_parser.h"
#include "performance_calculate.h"
#include "spmv_b200.h"
#include "utility.h"

#define DEFAULT_ITERS 95 /* reference NUM_ITERATION - ITERATION_SKIP, main_cuda.cu:20 */

typedef struct {
    const char *name;
    double time, min_time, flops, gbs, roofline;
    DiffMetrics err;
    int ran;
} KernelRow;

enum { K_CSR_AUTO, K_CSR_ROW, K_CSR_STREAM, K_CSR_VECTOR, K_CSR_BINNED, K_HLL_AUTO, K_HLL_ROWS, K_HLL_STREAM, K_HLL_SLICE, K_COUNT };
static const char *kNames[K_COUNT] = {"csr_auto", "csr_row", "csr_stream", "csr_vector", "csr_binned", "hll_auto", "hll_rows", "hll_stream", "hll_slice"};

typedef struct {
    const char *csv;
    int iters, warmup;
    double peak_gbs;
    int gpus;
} Options;

typedef struct {
    int ngpus;              /* GPUs the partitioner actually used */
    double time, flops, lambda, rel_err;
} MultiRow;

static int die_cuda(const char *what) {
    fprintf(stderr, "spmv_driver: %s: %s\n", what, spmv_b200_last_error());
    return 3;
}

static void write_csv(const Options *opt, const char *matrix, int M, int N, long long nz, const KernelRow *k,
                      double e2e_csr, double e2e_hll, const MultiRow *multi) {
    FILE *fp = fopen(opt->csv, "a+");
    if (!fp) {
        printf("Errore nell'apertura del file %s\n", opt->csv);
        return;
    }
    fseek(fp, 0, SEEK_END);
    if (ftell(fp) == 0) { /* header only for a new file, as the reference does */
        fprintf(fp, "matrix_name,rows,cols,nonzeros");
        for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",time_%s", kNames[i]);
        for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",flops_%s", kNames[i]);
        for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",gbs_%s", kNames[i]);
        for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",roofline_%s", kNames[i]);
        for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",relative_error_%s,absolute_error_%s", kNames[i], kNames[i]);
        fprintf(fp, ",time_e2e_csr_host,time_e2e_hll_host,peak_gbs,ngpus,check_baseline");
        fprintf(fp, ",time_multi_iteration,flops_multi_iteration,lambda_multi,relative_error_multi_product\n");
    }
    fprintf(fp, "%s,%d,%d,%lld", matrix, M, N, nz);
    for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",%.15f", k[i].time);
    for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",%.15f", k[i].flops);
    for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",%.6f", k[i].gbs);
    for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",%.6f", k[i].roofline);
    for (int i = 0; i < K_COUNT; ++i) fprintf(fp, ",%.15f,%.15f", k[i].err.mean_rel_err, k[i].err.mean_abs_err);
    /* the error columns compare with this library's one-thread-per-row kernel in the reference's summation order (bit-identical
     * to csr_matrix_vector_mult, tests/test_gpu_parity.py) -- NOT with a CPU product: the library has no CPU path */
    fprintf(fp, ",%.15f,%.15f,%.1f,%d,gpu_serial_order_kernel", e2e_csr, e2e_hll, opt->peak_gbs, multi ? multi->ngpus : 1);
    fprintf(fp, ",%.15f,%.15f,%.15g,%.15g\n", multi ? multi->time : 0.0, multi ? multi->flops : 0.0, multi ? multi->lambda : 0.0,
            multi ? multi->rel_err : 0.0);
    fclose(fp);
}

/* everything after the matrix is resident: timing, checks, report */
static int run_resident(const Options *opt, const char *name, spmv_b200_csr *A, spmv_b200_hll *H, const CSRMatrix *host) {
    spmv_b200_csr_info_t ci;
    spmv_b200_hll_info_t hi;
    if (spmv_b200_csr_info(A, &ci) || spmv_b200_hll_info(H, &hi)) return die_cuda("info");
    const int M = ci.M, N = ci.N;
    double *x = malloc((size_t)(N > 0 ? N : 1) * sizeof(double));
    double *y_ref = calloc((size_t)(M > 0 ? M : 1), sizeof(double));
    double *y = calloc((size_t)(M > 0 ? M : 1), sizeof(double));
    if (!x || !y_ref || !y) {
        fprintf(stderr, "spmv_driver: out of host memory\n");
        free(x), free(y_ref), free(y);
        return 1;
    }
    init_vector_at_one(x, N); /* the reference's x (main_cuda.cu:76, src/utility.c:18-22) */

    /* check baseline: one thread per row, left to right -- the summation order of csr_matrix_vector_mult */
    int rc = spmv_b200_csr_replan(A, 0, 0, 1, NULL);
    if (!rc) rc = spmv_b200_csr_spmv_host(A, x, y_ref, 0, SPMV_B200_ALGO_AUTO);
    if (!rc) rc = spmv_b200_csr_replan(A, 0, 0, 0, NULL);
    if (rc) {
        free(x), free(y_ref), free(y);
        return die_cuda("serial-order product");
    }

    KernelRow k[K_COUNT];
    memset(k, 0, sizeof k);
    const int csr_algo[K_HLL_AUTO] = {SPMV_B200_ALGO_AUTO, SPMV_B200_ALGO_ROW, SPMV_B200_ALGO_STREAM, SPMV_B200_ALGO_VECTOR, SPMV_B200_ALGO_BINNED};
    const int hll_kernel[K_COUNT - K_HLL_AUTO] = {0, 3, 2, 1};
    int bad = 0;
    printf("\n=== %s: %d x %d, %lld nonzeros, %d hacks, %lld HLL slots ===\n", name, M, N, ci.nnz, hi.num_hacks, hi.slots);
    for (int i = 0; i < K_COUNT; ++i) {
        const int is_hll = i >= K_HLL_AUTO;
        const long long bytes = is_hll ? hi.algorithmic_bytes : ci.algorithmic_bytes;
        k[i].name = kNames[i];
        memset(y, 0, (size_t)M * sizeof(double));
        rc = is_hll ? spmv_b200_hll_time(H, x, y, hll_kernel[i - K_HLL_AUTO], opt->warmup, opt->iters, &k[i].time, &k[i].min_time)
                    : spmv_b200_csr_time(A, x, y, csr_algo[i], opt->warmup, opt->iters, &k[i].time, &k[i].min_time);
        if (rc) {
            free(x), free(y_ref), free(y);
            return die_cuda(kNames[i]);
        }
        k[i].ran = 1;
        k[i].flops = calculate_flops((int)ci.nnz, k[i].time);
        k[i].gbs = calculate_bandwidth_gbs(bytes, k[i].time);
        k[i].roofline = calculate_roofline_fraction(bytes, k[i].time, opt->peak_gbs);
        k[i].err = computeDifferenceMetrics(y_ref, y, M, 1e-5, 1e-4, false); /* the reference's tolerances, main.c:145-362 */
        if (k[i].err.significant_diffs) bad = 1;
        printf("%-11s mean %.6f s  min %.6f s  ", kNames[i], k[i].time, k[i].min_time);
        print_flops(k[i].flops);
        printf("            %.1f GB/s = %.1f %% of %.0f GB/s   rel.err %.3g  abs.err %.3g  diffs %d\n", k[i].gbs,
               100.0 * k[i].roofline, opt->peak_gbs, k[i].err.mean_rel_err, k[i].err.mean_abs_err, k[i].err.significant_diffs);
    }

    /* host-buffer (end-to-end) products: H2D x, product, D2H y in one pipelined call */
    double e2e[2] = {0.0, 0.0};
    for (int f = 0; f < 2; ++f) {
        const MediumPerformanceMetric id = f ? B200_HLL_E2E_TIME : B200_CSR_E2E_TIME;
        reset_medium_time_metrics();
        for (int it = 0; it < opt->warmup + 5; ++it) {
            const double t0 = omp_get_wtime();
            rc = f ? spmv_b200_hll_spmv_host(H, x, y) : spmv_b200_csr_spmv_host(A, x, y, 0, SPMV_B200_ALGO_AUTO);
            const double t1 = omp_get_wtime();
            if (rc) {
                free(x), free(y_ref), free(y);
                return die_cuda("host product");
            }
            if (it >= opt->warmup) update_medium_metric(id, t1 - t0);
        }
        e2e[f] = get_metric_value(id);
        DiffMetrics d = computeDifferenceMetrics(y_ref, y, M, 1e-5, 1e-4, false);
        if (d.significant_diffs) bad = 1;
        printf("%-11s mean %.6f s (host x -> device, product, device -> host y)  diffs %d\n", f ? "hll_host" : "csr_host", e2e[f],
               d.significant_diffs);
    }
    /* N GPUs of this box through the single-process entry points (spmv_b200_multi_*): the row-partitioned product against
     * the same serial-order baseline, then the power method (y = A x; lambda = |y|; x = y / lambda) */
    MultiRow multi = {1, 0.0, 0.0, 0.0, 0.0};
    int have_multi = 0;
    if (opt->gpus >= 1 && host && M == N && M > 0) {
        spmv_b200_multi *ctx = NULL;
        spmv_b200_multi_info_t mi;
        rc = spmv_b200_multi_init_csr(opt->gpus, SPMV_B200_FORMAT_CSR, host->M, host->N, host->nz, host->row_ptr, host->col_idx,
                                      host->values, &ctx);
        if (!rc) rc = spmv_b200_multi_info(ctx, &mi);
        if (!rc) rc = spmv_b200_multi_spmv(ctx, x, y);
        if (!rc) {
            DiffMetrics d = computeDifferenceMetrics(y_ref, y, M, 1e-5, 1e-4, false);
            if (d.significant_diffs) bad = 1;
            multi.rel_err = d.mean_rel_err;
            multi.ngpus = mi.ngpus;
            const int mode = mi.fused_ok ? SPMV_B200_EXCHANGE_MAILBOX : SPMV_B200_EXCHANGE_ALLGATHER_PEER;
            double ms = 0.0;
            rc = spmv_b200_multi_reset(ctx, NULL);
            if (!rc) rc = spmv_b200_multi_iterate(ctx, 3, mode, NULL, NULL); /* warm-up */
            if (!rc) rc = spmv_b200_multi_iterate(ctx, 20, mode, &multi.lambda, &ms);
            multi.time = ms * 1e-3;
            multi.flops = calculate_flops((int)ci.nnz, multi.time);
            have_multi = !rc;
            if (!rc)
                printf("%d GPU(s)    product rel.err %.3g diffs %d; power iteration (%s) %.6f s / iteration, lambda %.12g  ", mi.ngpus,
                       d.mean_rel_err, d.significant_diffs, mi.fused_ok ? "fused mailbox exchange" : "peer all-gather", multi.time, multi.lambda);
            if (!rc) print_flops(multi.flops);
        }
        spmv_b200_multi_free(ctx);
        if (rc) {
            free(x), free(y_ref), free(y);
            return die_cuda("multi-GPU run");
        }
    }
    if (opt->csv) write_csv(opt, name, M, N, ci.nnz, k, e2e[0], e2e[1], have_multi ? &multi : NULL);
    free(x), free(y_ref), free(y);
    return bad ? 2 : 0;
}

static int run_file(const Options *opt, const char *path) {
    PreMatrix pre;
    CSRMatrix csr;
    HLLMatrix hll;
    init_pre_matrix(&pre);
    init_csr_matrix(&csr);
    init_hll_matrix(&hll);
    const char *base = strrchr(path, '/');
    base = base ? base + 1 : path;
    int status = 1;
    spmv_b200_csr *A = NULL;
    spmv_b200_hll *H = NULL;
    if (read_matrix_market(path, &pre) != 0) goto out;
    if (convert_in_csr(&pre, &csr, base) != 0) goto out;
    if (convert_to_hll(&pre, &hll) != 0) goto out;
    if (spmv_b200_csr_upload(csr.M, csr.N, csr.nz, csr.row_ptr, csr.col_idx, csr.values, &A) ||
        spmv_b200_hll_upload(&hll, csr.M, csr.N, &H)) {
        status = die_cuda("upload");
        goto out;
    }
    status = run_resident(opt, base, A, H, &csr);
out:
    spmv_b200_hll_free(H);
    spmv_b200_csr_free(A);
    free_hll_matrix(&hll);
    free_csr_matrix(&csr);
    free_pre_matrix(&pre);
    return status;
}

static int run_lap2d(const Options *opt, int n) {
    spmv_b200_csr *A = NULL;
    spmv_b200_hll *H = NULL;
    char name[64];
    snprintf(name, sizeof name, "lap2d_%d", n);
    int status;
    if (spmv_b200_synth_csr(SPMV_B200_SYNTH_LAP2D, n, 0, 0, 0, 0, (long long)n * n, NULL, &A) ||
        spmv_b200_hll_from_csr(A, NULL, &H))
        status = die_cuda("synthetic matrix");
    else
        status = run_resident(opt, name, A, H, NULL);
    spmv_b200_hll_free(H);
    spmv_b200_csr_free(A);
    return status;
}

int main(int argc, char **argv) {
    Options opt = {NULL, DEFAULT_ITERS, ITERATION_SKIP, 8000.0, 0};
    int lap2d[8], nlap = 0, status = 0, ran = 0;
    initialize_metrics();
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--csv") && i + 1 < argc) opt.csv = argv[++i];
        else if (!strcmp(argv[i], "--iters") && i + 1 < argc) opt.iters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) opt.warmup = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--peak-gbs") && i + 1 < argc) opt.peak_gbs = atof(argv[++i]);
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) opt.gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--lap2d") && i + 1 < argc && nlap < 8) lap2d[nlap++] = atoi(argv[++i]);
        else if (argv[i][0] == '-') {
            fprintf(stderr, "usage: %s [--csv out.csv] [--iters 95] [--warmup 5] [--peak-gbs 8000] [--gpus N] [--lap2d n] matrix.mtx ...\n", argv[0]);
            return 1;
        }
    }
    if (opt.iters <= 0 || opt.warmup < 0) {
        fprintf(stderr, "spmv_driver: --iters must be positive, --warmup non-negative\n");
        return 1;
    }
    int devices = 0;
    if (spmv_b200_device_count(&devices) != SPMV_B200_OK || devices == 0) {
        fprintf(stderr, "spmv_driver: no CUDA device (%s); this engine has no CPU fallback\n", spmv_b200_last_error());
        cleanup_metrics();
        return 3;
    }
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--csv") || !strcmp(argv[i], "--iters") || !strcmp(argv[i], "--warmup") ||
            !strcmp(argv[i], "--peak-gbs") || !strcmp(argv[i], "--lap2d") || !strcmp(argv[i], "--gpus")) {
            ++i;
            continue;
        }
        const int rc = run_file(&opt, argv[i]);
        if (rc > status) status = rc;
        ++ran;
    }
    for (int i = 0; i < nlap; ++i) {
        const int rc = run_lap2d(&opt, lap2d[i]);
        if (rc > status) status = rc;
        ++ran;
    }
    cleanup_metrics();
    if (!ran) {
        fprintf(stderr, "spmv_driver: nothing to do (give .mtx files or --lap2d n)\n");
        return 1;
    }
    return status;
}
