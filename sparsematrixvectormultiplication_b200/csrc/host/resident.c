/*
 * resident.c -- opt-in cache of device copies behind the reference's stateless product signatures.
 * One CSR and one HLL entry (the reference's drivers work on one matrix at a time, main.c:56-470).
 */
#include "resident.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int g_enabled = -1; /* -1: not decided yet (environment) */

/* 0 off, 1 sampled fingerprint (the documented contract: call spmv_b200_resident_drop() after changing the arrays in
 * place), 2 strict: every element of the index and value arrays is hashed on every call, so an in-place change or a
 * free + malloc at the same addresses can never go unnoticed (costs one pass over the host arrays per product). */
static int enabled(void) {
    if (g_enabled < 0) {
        const char *v = getenv("SPMV_B200_RESIDENT");
        g_enabled = (v && *v && strcmp(v, "0") != 0) ? (strcmp(v, "2") == 0 ? 2 : 1) : 0;
    }
    return g_enabled;
}

/* 64 evenly spaced samples of the index and value arrays, mixed FNV-style */
static uint64_t mix(uint64_t h, uint64_t v) { return (h ^ v) * 0x100000001B3ULL; }

static uint64_t hash_span(const int *idx, const double *val, long long lo, long long hi) {
    uint64_t h = 0xCBF29CE484222325ULL;
    for (long long k = lo; k < hi; ++k) {
        uint64_t bits;
        memcpy(&bits, &val[k], sizeof bits);
        h = mix(mix(h, (uint64_t)(uint32_t)idx[k]), bits);
    }
    return h;
}

static uint64_t sample_arrays(const int *idx, const double *val, long long n) {
    uint64_t h = 0xCBF29CE484222325ULL;
    if (n <= 0) return h;
    if (enabled() == 2) { /* strict: all of it, in fixed 1 Mi-element spans hashed in parallel and chained in order */
        const long long span = 1LL << 20, spans = (n + span - 1) / span;
        uint64_t *part = malloc((size_t)spans * sizeof *part);
        if (part) {
#pragma omp parallel for schedule(static)
            for (long long c = 0; c < spans; ++c) part[c] = hash_span(idx, val, c * span, (c + 1) * span < n ? (c + 1) * span : n);
            for (long long c = 0; c < spans; ++c) h = mix(h, part[c]);
            free(part);
            return h;
        }
        return hash_span(idx, val, 0, n);
    }
    const long long step = n > 64 ? n / 64 : 1;
    for (long long k = 0; k < n; k += step) {
        uint64_t bits;
        memcpy(&bits, &val[k], sizeof bits);
        h = mix(mix(h, (uint64_t)(uint32_t)idx[k]), bits);
    }
    uint64_t last;
    memcpy(&last, &val[n - 1], sizeof last);
    return mix(mix(h, (uint64_t)(uint32_t)idx[n - 1]), last);
}

static struct {
    spmv_b200_csr *handle;
    const int *row_ptr, *col_idx;
    const double *values;
    int M, N;
    long long nnz;
    uint64_t print;
} g_csr;

static struct {
    spmv_b200_hll *handle;
    const ELLPACKBlock *blocks;
    int count, rows, N;
    uint64_t print;
} g_hll;

int spmv_b200_resident_cache(int enable) {
    const int before = enabled();
    g_enabled = enable ? (enable == 2 ? 2 : 1) : 0;
    if (!g_enabled) spmv_b200_resident_drop();
    return before;
}

void spmv_b200_resident_drop(void) {
    spmv_b200_csr_free(g_csr.handle);
    memset(&g_csr, 0, sizeof g_csr);
    spmv_b200_hll_free(g_hll.handle);
    memset(&g_hll, 0, sizeof g_hll);
}

int resident_csr(int M, int N, long long nnz, const int *row_ptr, const int *col_idx, const double *values,
                 spmv_b200_csr **out, int *borrowed) {
    *borrowed = 0;
    if (!enabled()) { /* used once: timing kernel candidates at upload would cost more than it can save */
        const int tune = spmv_b200_autotune(0);
        const int rc = spmv_b200_csr_upload(M, N, nnz, row_ptr, col_idx, values, out);
        spmv_b200_autotune(tune);
        return rc;
    }
    const uint64_t print = mix(sample_arrays(col_idx, values, nnz), (uint64_t)(uint32_t)row_ptr[M]);
    if (g_csr.handle && g_csr.row_ptr == row_ptr && g_csr.col_idx == col_idx && g_csr.values == values && g_csr.M == M &&
        g_csr.N == N && g_csr.nnz == nnz && g_csr.print == print) {
        *out = g_csr.handle;
        *borrowed = 1;
        return SPMV_B200_OK;
    }
    spmv_b200_csr_free(g_csr.handle);
    memset(&g_csr, 0, sizeof g_csr);
    const int rc = spmv_b200_csr_upload(M, N, nnz, row_ptr, col_idx, values, out);
    if (rc != SPMV_B200_OK) return rc;
    g_csr.handle = *out;
    g_csr.row_ptr = row_ptr;
    g_csr.col_idx = col_idx;
    g_csr.values = values;
    g_csr.M = M;
    g_csr.N = N;
    g_csr.nnz = nnz;
    g_csr.print = print;
    *borrowed = 1;
    return SPMV_B200_OK;
}

static uint64_t sample_blocks(const ELLPACKBlock *blocks, int count) {
    uint64_t h = 0xCBF29CE484222325ULL;
    const int step = (count > 16 && enabled() != 2) ? count / 16 : 1;
    for (int b = 0; b < count; b += step) {
        const ELLPACKBlock *blk = &blocks[b];
        h = mix(mix(mix(h, (uint64_t)blk->M), (uint64_t)blk->MAXNZ), (uint64_t)(uintptr_t)blk->JA);
        if (blk->MAXNZ > 0 && blk->JA && blk->AS) h = mix(h, sample_arrays(blk->JA, blk->AS, (long long)blk->M * blk->MAXNZ));
    }
    return h;
}

int resident_hll(const ELLPACKBlock *blocks, int count, int rows, int N, spmv_b200_hll **out, int *borrowed) {
    *borrowed = 0;
    HLLMatrix view = {count, (ELLPACKBlock *)blocks};
    if (!enabled()) {
        const int tune = spmv_b200_autotune(0);
        const int rc = spmv_b200_hll_upload(&view, rows, N, out);
        spmv_b200_autotune(tune);
        return rc;
    }
    const uint64_t print = sample_blocks(blocks, count);
    if (g_hll.handle && g_hll.blocks == blocks && g_hll.count == count && g_hll.rows == rows && g_hll.N == N &&
        g_hll.print == print) {
        *out = g_hll.handle;
        *borrowed = 1;
        return SPMV_B200_OK;
    }
    spmv_b200_hll_free(g_hll.handle);
    memset(&g_hll, 0, sizeof g_hll);
    const int rc = spmv_b200_hll_upload(&view, rows, N, out);
    if (rc != SPMV_B200_OK) return rc;
    g_hll.handle = *out;
    g_hll.blocks = blocks;
    g_hll.count = count;
    g_hll.rows = rows;
    g_hll.N = N;
    g_hll.print = print;
    *borrowed = 1;
    return SPMV_B200_OK;
}

void resident_forget_csr(const void *p) {
    if (g_csr.handle && p && (p == g_csr.row_ptr || p == g_csr.col_idx || p == g_csr.values)) {
        spmv_b200_csr_free(g_csr.handle);
        memset(&g_csr, 0, sizeof g_csr);
    }
}

void resident_forget_hll(const void *blocks) {
    if (g_hll.handle && blocks && blocks == (const void *)g_hll.blocks) {
        spmv_b200_hll_free(g_hll.handle);
        memset(&g_hll, 0, sizeof g_hll);
    }
}
