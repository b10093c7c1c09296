/* multi_check.c -- drives the single-process multi-GPU entry points of include/spmv_b200.h from plain C.
 *
 *   multi_check <ngpus> [n]
 *
 * For the n^3 7-point Laplacian (default n = 40): the power iteration on 1 GPU and on <ngpus> GPUs, CSR and HLL, with the
 * fused MAILBOX exchange and with the literal NCCL ALLGATHER exchange, must give the same lambda (1e-12 relative: only
 * the association of the N-GPU sum of |y|^2 differs) and the same normalised iterate; the row-partitioned product on
 * host vectors must be bitwise the single-GPU product; a host CSR matrix (spmv_b200_multi_init_csr) must behave like
 * the generated one.  Prints one "lambda ..." line per run (the pytest wrapper compares them with the CPU oracle) and
 * MULTI_CHECK OK / FAILED.  Test infrastructure: not part of the product. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "spmv_b200.h"

#define ITERS 12
static int failures = 0;

#define REQUIRE(cond, ...)                      \
    do {                                        \
        if (!(cond)) {                          \
            printf("FAIL: " __VA_ARGS__);       \
            printf("\n");                       \
            ++failures;                         \
        }                                       \
    } while (0)

#define CALL(expr)                                                              \
    do {                                                                        \
        if ((expr) != SPMV_B200_OK) {                                           \
            printf("FAIL: %s: %s\n", #expr, spmv_b200_last_error());            \
            return 1;                                                           \
        }                                                                       \
    } while (0)

static double max_abs_diff(const double *a, const double *b, long long n) {
    double worst = 0.0;
    for (long long i = 0; i < n; ++i) {
        const double d = fabs(a[i] - b[i]);
        if (d > worst || d != d) worst = d;
    }
    return worst;
}

/* one run: reset, ITERS iterations in two calls (the iterate continues across calls), lambda and x */
static int run(spmv_b200_multi *ctx, int exchange, const char *label, double *lambda, double *x) {
    double ms = 0.0, lam_half = 0.0;
    CALL(spmv_b200_multi_reset(ctx, NULL));
    CALL(spmv_b200_multi_iterate(ctx, ITERS / 2, exchange, &lam_half, NULL));
    CALL(spmv_b200_multi_iterate(ctx, ITERS - ITERS / 2, exchange, lambda, &ms));
    CALL(spmv_b200_multi_get_x(ctx, x));
    spmv_b200_multi_info_t info;
    CALL(spmv_b200_multi_info(ctx, &info));
    printf("lambda %-28s gpus %d iters %d: %.17g  (%.4f ms/iteration, halo of GPU 0: %lld doubles)\n", label, info.ngpus, ITERS, *lambda, ms,
           info.halo_doubles[0]);
    return 0;
}

int main(int argc, char **argv) {
    const int ngpus = argc > 1 ? atoi(argv[1]) : 1;
    const int n = argc > 2 ? atoi(argv[2]) : 40;
    const long long M = (long long)n * n * n;
    int visible = 0;
    CALL(spmv_b200_device_count(&visible));
    printf("multi_check: %d GPUs requested, %d visible, lap3d %d^3\n", ngpus, visible, n);
    double *x_ref = malloc((size_t)M * sizeof(double)), *x = malloc((size_t)M * sizeof(double));
    double *xin = malloc((size_t)M * sizeof(double)), *y1 = malloc((size_t)M * sizeof(double)), *yn = malloc((size_t)M * sizeof(double));
    if (!x_ref || !x || !xin || !y1 || !yn) return 2;
    for (long long i = 0; i < M; ++i) xin[i] = 1.0 + (double)(i % 7) / 8.0;

    /* reference: ONE GPU, CSR, fused */
    spmv_b200_multi *one = NULL;
    double lam_ref = 0.0;
    CALL(spmv_b200_multi_init_synth(1, SPMV_B200_FORMAT_CSR, SPMV_B200_SYNTH_LAP3D, n, 0, 0, 0x5EED, &one));
    if (run(one, SPMV_B200_EXCHANGE_MAILBOX, "csr/mailbox", &lam_ref, x_ref)) return 1;
    CALL(spmv_b200_multi_reset(one, NULL));
    CALL(spmv_b200_multi_spmv(one, xin, y1));

    const int formats[2] = {SPMV_B200_FORMAT_CSR, SPMV_B200_FORMAT_HLL};
    const char *fname[2] = {"csr", "hll"};
    const int modes[3] = {SPMV_B200_EXCHANGE_MAILBOX, SPMV_B200_EXCHANGE_ALLGATHER, SPMV_B200_EXCHANGE_ALLGATHER_PEER};
    const char *mname[3] = {"mailbox", "allgather", "allgather_peer"};
    for (int f = 0; f < 2; ++f) {
        spmv_b200_multi *ctx = NULL;
        CALL(spmv_b200_multi_init_synth(ngpus, formats[f], SPMV_B200_SYNTH_LAP3D, n, 0, 0, 0x5EED, &ctx));
        spmv_b200_multi_info_t info;
        CALL(spmv_b200_multi_info(ctx, &info));
        REQUIRE(info.M == M && info.nnz == 7 * M - 6LL * n * n, "info: M %lld nnz %lld", info.M, info.nnz);
        REQUIRE(info.row_begin[0] == 0 && info.row_end[info.ngpus - 1] == M, "partition does not cover the rows");
        for (int g = 1; g < info.ngpus; ++g) {
            REQUIRE(info.row_begin[g] == info.row_end[g - 1], "partition has a gap at GPU %d", g);
            if (formats[f] == SPMV_B200_FORMAT_HLL) REQUIRE(info.row_begin[g] % 32 == 0, "HLL cut %lld is not on a hack boundary", info.row_begin[g]);
        }
        for (int m = 0; m < 3; ++m) {
            char label[64];
            double lam = 0.0;
            snprintf(label, sizeof label, "%s/%s", fname[f], mname[m]);
            if (run(ctx, modes[m], label, &lam, x)) return 1;
            REQUIRE(fabs(lam - lam_ref) <= 1e-12 * lam_ref, "%s: lambda %.17g vs %.17g on one GPU", label, lam, lam_ref);
            REQUIRE(max_abs_diff(x, x_ref, M) <= 1e-12, "%s: x differs from the single-GPU iterate by %.3e", label, max_abs_diff(x, x_ref, M));
        }
        /* the exchange mode may not change without a reset */
        REQUIRE(spmv_b200_multi_iterate(ctx, 1, SPMV_B200_EXCHANGE_MAILBOX, NULL, NULL) == SPMV_B200_ERR_INVALID, "mode switch without reset accepted");
        CALL(spmv_b200_multi_reset(ctx, NULL));
        CALL(spmv_b200_multi_spmv(ctx, xin, yn));
        REQUIRE(memcmp(y1, yn, (size_t)M * sizeof(double)) == 0, "%s: the row-partitioned product is not bitwise the single-GPU product", fname[f]);
        /* a caller-supplied start vector */
        CALL(spmv_b200_multi_reset(ctx, xin));
        CALL(spmv_b200_multi_get_x(ctx, x));
        REQUIRE(memcmp(x, xin, (size_t)M * sizeof(double)) == 0, "%s: reset(x0) / get_x round trip", fname[f]);
        spmv_b200_multi_free(ctx);
    }

    /* host CSR arrays through the reference's own partitioner: same matrix, built here */
    {
        const long long n2 = (long long)n * n, nnz = 7 * M - 6 * n2;
        int *rp = malloc((size_t)(M + 1) * sizeof(int)), *ci = malloc((size_t)nnz * sizeof(int));
        double *va = malloc((size_t)nnz * sizeof(double));
        if (!rp || !ci || !va) return 2;
        long long at = 0;
        rp[0] = 0;
        for (long long r = 0; r < M; ++r) {
            const long long i = r / n2, j = (r / n) % n, k = r % n;
            const long long cand[7] = {r - n2, r - n, r - 1, r, r + 1, r + n, r + n2};
            const int on[7] = {i > 0, j > 0, k > 0, 1, k < n - 1, j < n - 1, i < n - 1};
            for (int e = 0; e < 7; ++e)
                if (on[e]) {
                    ci[at] = (int)cand[e];
                    va[at++] = e == 3 ? 6.0 : -1.0;
                }
            rp[r + 1] = (int)at;
        }
        spmv_b200_multi *ctx = NULL;
        double lam = 0.0;
        CALL(spmv_b200_multi_init_csr(ngpus, SPMV_B200_FORMAT_CSR, (int)M, (int)M, nnz, rp, ci, va, &ctx));
        if (run(ctx, SPMV_B200_EXCHANGE_MAILBOX, "host-csr/mailbox", &lam, x)) return 1;
        REQUIRE(fabs(lam - lam_ref) <= 1e-12 * lam_ref, "host CSR: lambda %.17g vs %.17g", lam, lam_ref);
        REQUIRE(max_abs_diff(x, x_ref, M) <= 1e-12, "host CSR: x differs by %.3e", max_abs_diff(x, x_ref, M));
        spmv_b200_multi_free(ctx);
        free(rp);
        free(ci);
        free(va);
    }
    spmv_b200_multi_free(one);
    free(x_ref);
    free(x);
    free(xin);
    free(y1);
    free(yn);
    printf(failures ? "MULTI_CHECK FAILED (%d)\n" : "MULTI_CHECK OK\n", failures);
    return failures ? 1 : 0;
}
