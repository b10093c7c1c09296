/*
 * oracle.c -- CPU ORACLE (test infrastructure, NOT the product). See oracle.h.
 *
 * Restates, in plain C, what the reference does on its CSR / HLL SpMV path.  Written from the
 * behaviour described in SURVEY.md section 3.3 / 3.4 and pinned against the real reference
 * (oracle/_ref) and tests/golden/.  Parity: PINNED (see tests/test_oracle_pinned.py).
 */
#include "oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Matrix Market reader.  reference: libs/mmio.c:96-178 (banner), :189-214 (size line),
 * src/matrix_parser.c:25-150 (entries, 1->0 based, bounds, symmetric mirroring, pattern=1.0).
 * ---------------------------------------------------------------------------------------- */
static void lower(char *s) { for (; *s; ++s) *s = (char)tolower((unsigned char)*s); }

static int banner(FILE *f, char type[4]) {
    char line[1025], tok[5][64];
    type[0] = type[1] = type[2] = ' ';
    type[3] = 'G';
    if (!fgets(line, sizeof line, f)) return -1;
    if (sscanf(line, "%63s %63s %63s %63s %63s", tok[0], tok[1], tok[2], tok[3], tok[4]) != 5)
        return -1;
    for (int k = 1; k < 5; ++k) lower(tok[k]);
    if (strncmp(tok[0], "%%MatrixMarket", 14) != 0) return -1;
    if (strcmp(tok[1], "matrix") != 0) return -1;
    type[0] = 'M';
    if (!strcmp(tok[2], "coordinate")) type[1] = 'C';
    else if (!strcmp(tok[2], "array")) type[1] = 'A';
    else return -1;
    if (!strcmp(tok[3], "real")) type[2] = 'R';
    else if (!strcmp(tok[3], "complex")) type[2] = 'C';
    else if (!strcmp(tok[3], "pattern")) type[2] = 'P';
    else if (!strcmp(tok[3], "integer")) type[2] = 'I';
    else return -1;
    if (!strcmp(tok[4], "general")) type[3] = 'G';
    else if (!strcmp(tok[4], "symmetric")) type[3] = 'S';
    else if (!strcmp(tok[4], "hermitian")) type[3] = 'H';
    else if (!strcmp(tok[4], "skew-symmetric")) type[3] = 'K';
    else return -1;
    return 0;
}

static int size_line(FILE *f, int *M, int *N, int *nz) {
    char line[1025];
    *M = *N = *nz = 0;
    do {
        if (!fgets(line, sizeof line, f)) return -1;
    } while (line[0] == '%');
    if (sscanf(line, "%d %d %d", M, N, nz) == 3) return 0;
    for (;;) { /* blank line(s) before the size line */
        int got = fscanf(f, "%d %d %d", M, N, nz);
        if (got == EOF) return -1;
        if (got == 3) return 0;
    }
}

void orc_free_coo(orc_coo *c) {
    if (!c) return;
    free(c->I); free(c->J); free(c->val);
    memset(c, 0, sizeof *c);
}

int orc_read_matrix_market(const char *path, orc_coo *out) {
    memset(out, 0, sizeof *out);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int declared = 0;
    if (banner(f, out->type) != 0 || out->type[0] != 'M' || out->type[1] != 'C' ||
        size_line(f, &out->M, &out->N, &declared) != 0) {
        fclose(f);
        return -1;
    }
    const int sym = out->type[3] == 'S';     /* only 'S' is mirrored: matrix_parser.c:53,116 */
    const int pattern = out->type[2] == 'P';
    const long long cap = sym ? 2LL * declared : declared;
    int *I = malloc((size_t)(cap ? cap : 1) * sizeof(int));
    int *J = malloc((size_t)(cap ? cap : 1) * sizeof(int));
    double *V = malloc((size_t)(cap ? cap : 1) * sizeof(double));
    if (!I || !J || !V) { free(I); free(J); free(V); fclose(f); return -1; }
    int n = 0, bad = 0;
    for (int e = 0; e < declared && !bad; ++e) {
        int r, c;
        double v = 1.0;
        if (pattern) { if (fscanf(f, "%d %d", &r, &c) != 2) bad = 1; }
        else if (fscanf(f, "%d %d %lf", &r, &c, &v) != 3) bad = 1;
        if (bad) break;
        --r; --c;
        if (r < 0 || r >= out->M || c < 0 || c >= out->N) { bad = 1; break; }
        I[n] = r; J[n] = c; V[n] = v; ++n;
        if (sym && r != c) { I[n] = c; J[n] = r; V[n] = v; ++n; } /* mirror right behind it */
    }
    fclose(f);
    if (bad) { free(I); free(J); free(V); return -1; }
    out->nz = n; out->I = I; out->J = J; out->val = V;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * COO -> CSR.  reference src/csr_matrix.c:63-126; row sort = src/utility.c:38-91 (Lomuto
 * partition, pivot = last element, predicate col <= pivot; unstable, duplicates kept).
 * The order in which sub-ranges are visited (recursion vs the explicit stack used above
 * 10 000 elements) does not change the result: sub-ranges are disjoint.
 * ---------------------------------------------------------------------------------------- */
static long lomuto(int *c, double *v, long lo, long hi) {
    const int pivot = c[hi];
    long store = lo;
    for (long k = lo; k < hi; ++k) {
        if (c[k] <= pivot) {
            int tc = c[store]; c[store] = c[k]; c[k] = tc;
            double tv = v[store]; v[store] = v[k]; v[k] = tv;
            ++store;
        }
    }
    int tc = c[store]; c[store] = c[hi]; c[hi] = tc;
    double tv = v[store]; v[store] = v[hi]; v[hi] = tv;
    return store;
}

static void quicksort_row(int *c, double *v, long lo, long hi) {
    /* explicit, growable stack: equivalent result to the reference's recursion */
    long cap = 128, top = 0;
    long *st = malloc((size_t)cap * sizeof(long));
    st[top++] = lo; st[top++] = hi;
    while (top > 0) {
        hi = st[--top]; lo = st[--top];
        if (lo >= hi) continue;
        long p = lomuto(c, v, lo, hi);
        if (top + 4 > cap) { cap *= 2; st = realloc(st, (size_t)cap * sizeof(long)); }
        st[top++] = lo; st[top++] = p - 1;
        st[top++] = p + 1; st[top++] = hi;
    }
    free(st);
}

int orc_coo_to_csr(int M, int nz, const int *I, const int *J, const double *val,
                   int *row_ptr, int *col_idx, double *values) {
    memset(row_ptr, 0, (size_t)(M + 1) * sizeof(int));
    for (int k = 0; k < nz; ++k) row_ptr[I[k] + 1]++;
    for (int r = 0; r < M; ++r) row_ptr[r + 1] += row_ptr[r];
    int *cursor = malloc((size_t)(M ? M : 1) * sizeof(int));
    if (!cursor) return -1;
    memcpy(cursor, row_ptr, (size_t)M * sizeof(int));
    for (int k = 0; k < nz; ++k) { /* file order inside each row */
        int d = cursor[I[k]]++;
        col_idx[d] = J[k];
        values[d] = val[k];
    }
    free(cursor);
    for (int r = 0; r < M; ++r)
        if (row_ptr[r + 1] - row_ptr[r] > 1)
            quicksort_row(col_idx, values, row_ptr[r], (long)row_ptr[r + 1] - 1);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * COO -> HLL.  reference src/hll_matrix.c:37-257: blocks of 32 rows (last one short),
 * MAXNZ = longest row of the block, row-major slots r*MAXNZ + j, rows sorted by column with
 * libc qsort (glibc 2.39: merge sort => stable, duplicates stay in file order), padding
 * JA = last real column of the row (0 for an empty row), AS = 0.0.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int col; double val; } cv;

static void stable_sort_cv(cv *a, cv *tmp, int n) {
    if (n < 2) return;
    int h = n / 2;
    stable_sort_cv(a, tmp, h);
    stable_sort_cv(a + h, tmp, n - h);
    int i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].col < a[i].col) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, (size_t)n * sizeof(cv));
}

void orc_free_hll(orc_hll *h) {
    if (!h) return;
    free(h->rows); free(h->maxnz); free(h->offset); free(h->JA); free(h->AS);
    memset(h, 0, sizeof *h);
}

int orc_coo_to_hll(int M, int N, int nz, const int *I, const int *J, const double *val,
                   orc_hll *out) {
    (void)N;
    memset(out, 0, sizeof *out);
    const int nb = (M + ORC_HACK_SIZE - 1) / ORC_HACK_SIZE;
    int *cnt = calloc((size_t)(M ? M : 1), sizeof(int));
    long long *rstart = malloc((size_t)(M + 1) * sizeof(long long));
    cv *pairs = malloc((size_t)(nz ? nz : 1) * sizeof(cv));
    cv *tmp = malloc((size_t)(nz ? nz : 1) * sizeof(cv));
    out->num_blocks = nb;
    out->rows = malloc((size_t)(nb ? nb : 1) * sizeof(int));
    out->maxnz = malloc((size_t)(nb ? nb : 1) * sizeof(int));
    out->offset = malloc((size_t)(nb + 1) * sizeof(long long));
    if (!cnt || !rstart || !pairs || !tmp || !out->rows || !out->maxnz || !out->offset) return -1;
    for (int k = 0; k < nz; ++k) {
        if (I[k] < 0 || I[k] >= M) { free(cnt); free(rstart); free(pairs); free(tmp); return -1; }
        cnt[I[k]]++;
    }
    rstart[0] = 0;
    for (int r = 0; r < M; ++r) rstart[r + 1] = rstart[r] + cnt[r];
    long long total = 0;
    for (int b = 0; b < nb; ++b) {
        int r0 = b * ORC_HACK_SIZE, r1 = (b == nb - 1) ? M : r0 + ORC_HACK_SIZE, w = 0;
        for (int r = r0; r < r1; ++r) if (cnt[r] > w) w = cnt[r];
        out->rows[b] = r1 - r0;
        out->maxnz[b] = w;
        out->offset[b] = total;
        total += (long long)w * (r1 - r0);
    }
    out->offset[nb] = total;
    out->JA = calloc((size_t)(total ? total : 1), sizeof(int));
    out->AS = calloc((size_t)(total ? total : 1), sizeof(double));
    if (!out->JA || !out->AS) return -1;
    /* per-row gather in file order, then stable sort by column */
    int *fill = calloc((size_t)(M ? M : 1), sizeof(int));
    for (int k = 0; k < nz; ++k) {
        cv *p = &pairs[rstart[I[k]] + fill[I[k]]++];
        p->col = J[k]; p->val = val[k];
    }
    free(fill);
    for (int r = 0; r < M; ++r) stable_sort_cv(pairs + rstart[r], tmp, cnt[r]);
    for (int r = 0; r < M; ++r) {
        int b = r / ORC_HACK_SIZE, lr = r % ORC_HACK_SIZE, w = out->maxnz[b];
        int *ja = out->JA + out->offset[b] + (long long)lr * w;
        double *as = out->AS + out->offset[b] + (long long)lr * w;
        int last = 0;
        for (int j = 0; j < cnt[r]; ++j) {
            ja[j] = last = pairs[rstart[r] + j].col;
            as[j] = pairs[rstart[r] + j].val;
        }
        for (int j = cnt[r]; j < w; ++j) { ja[j] = last; as[j] = 0.0; }
    }
    free(cnt); free(rstart); free(pairs); free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Products.
 * ---------------------------------------------------------------------------------------- */
void orc_spmv_csr_serial(int M, const int *row_ptr, const int *col_idx, const double *values,
                         const double *x, double *y) {
    for (int r = 0; r < M; ++r)
        for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) y[r] += values[k] * x[col_idx[k]];
}

void orc_spmv_hll_serial(const orc_hll *h, const double *x, double *y) {
    for (int b = 0; b < h->num_blocks; ++b) {
        const int w = h->maxnz[b];
        for (int r = 0; r < h->rows[b]; ++r) {
            const int *ja = h->JA + h->offset[b] + (long long)r * w;
            const double *as = h->AS + h->offset[b] + (long long)r * w;
            double s = 0.0;
            for (int j = 0; j < w; ++j) s += as[j] * x[ja[j]];
            y[b * ORC_HACK_SIZE + r] = s;
        }
    }
}

/* Greedy split shared by both partitioners (reference src/csr_matrix.c:196-238 and
 * src/hll_matrix.c:464-511): walk the items, close the current part once its running weight
 * reaches ceil(total/T) (never the last part), then drop parts whose weight is 0.
 * The reference keeps the running weight in an int for rows and a long long for blocks;
 * both agree below 2^31. */
static int greedy_split(int n_items, const long long *weight, int T, long long total,
                        int *start, int *end) {
    if (n_items <= 0 || T <= 0) return 0;
    if (n_items < T) T = n_items;
    long long *w = calloc((size_t)T, sizeof(long long));
    for (int t = 0; t < T; ++t) start[t] = end[t] = -1;
    const long long target = (total + T - 1) / T;
    int cur = 0;
    long long run = 0;
    for (int i = 0; i < n_items; ++i) {
        if (start[cur] == -1) start[cur] = i;
        run += weight[i];
        w[cur] += weight[i];
        if (run >= target && cur < T - 1) { end[cur] = i + 1; ++cur; run = 0; }
    }
    if (cur < T) end[cur] = n_items;
    int used = 0;
    for (int t = 0; t < T; ++t)
        if (start[t] != -1 && end[t] != -1 && w[t] > 0) {
            start[used] = start[t]; end[used] = end[t]; ++used;
        }
    free(w);
    return used;
}

int orc_partition_rows(int M, const int *row_ptr, int num_threads, long long total_nnz,
                       int *start, int *end) {
    if (M <= 0 || num_threads <= 0) return 0;
    long long *w = malloc((size_t)M * sizeof(long long));
    for (int r = 0; r < M; ++r) w[r] = row_ptr[r + 1] - row_ptr[r];
    int used = greedy_split(M, w, num_threads, total_nnz, start, end);
    free(w);
    return used;
}

int orc_partition_hll_blocks(const orc_hll *h, int N, int num_threads, int *start, int *end) {
    if (!h || num_threads <= 0 || h->num_blocks <= 0) return 0;
    long long *w = malloc((size_t)h->num_blocks * sizeof(long long)), total = 0;
    for (int b = 0; b < h->num_blocks; ++b) {
        /* the reference counts slots whose JA lies in [0,N); it indexes them with the
         * column-major expression j*rows+i (src/hll_matrix.c:457) which visits every slot of
         * the block exactly once, so the count is order independent. */
        long long c = 0, slots = (long long)h->rows[b] * h->maxnz[b];
        for (long long s = 0; s < slots; ++s) {
            int col = h->JA[h->offset[b] + s];
            c += (col >= 0 && col < N);
        }
        w[b] = c;
        total += c;
    }
    int used = greedy_split(h->num_blocks, w, num_threads, total, start, end);
    free(w);
    return used;
}

void orc_spmv_csr_parallel(const int *row_ptr, const int *col_idx, const double *values,
                           const double *x, double *y, int num_threads,
                           const int *start, const int *end) {
#pragma omp parallel num_threads(num_threads)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        for (int r = start[t]; r < end[t]; ++r) {
            double s = 0.0;
            for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) s += values[k] * x[col_idx[k]];
            y[r] = s;
        }
    }
}

void orc_spmv_hll_parallel(const orc_hll *h, const double *x, double *y, int num_threads,
                           const int *start, const int *end) {
#pragma omp parallel num_threads(num_threads)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        for (int b = start[t]; b < end[t]; ++b) {
            const int w = h->maxnz[b];
            for (int r = 0; r < h->rows[b]; ++r) {
                const int *ja = h->JA + h->offset[b] + (long long)r * w;
                const double *as = h->AS + h->offset[b] + (long long)r * w;
                double s = 0.0;
                for (int j = 0; j < w; ++j) s += as[j] * x[ja[j]];
                y[b * ORC_HACK_SIZE + r] = s;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Harness arithmetic.
 * ---------------------------------------------------------------------------------------- */
double orc_calculate_flops(int nz, double seconds) { return 2.0 * nz / seconds; }

int orc_diff_metrics_c(const double *ref, const double *res, int n, double abs_tol,
                       double rel_tol, double *mean_rel_err) {
    int significant = 0;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        double d = fabs(ref[i] - res[i]);
        double den = fmax(fmax(fabs(ref[i]), fabs(res[i])), rel_tol);
        double rel = d > abs_tol ? d / den : 0.0;
        if (rel > rel_tol) { acc += rel; ++significant; }
    }
    *mean_rel_err = significant ? acc / significant : 0.0;
    return significant;
}

void orc_diff_metrics_cuda(const double *ref, const double *res, int n,
                           double *mean_abs_err, double *mean_rel_err) {
    double sa = 0.0, sr = 0.0;
    for (int i = 0; i < n; ++i) {
        double d = fabs(ref[i] - res[i]);
        double den = fmax(fmax(fabs(ref[i]), fabs(res[i])), 1e-4);
        sa += d;
        sr += d / den;
    }
    *mean_abs_err = n > 0 ? sa / n : 0.0;
    *mean_rel_err = n > 0 ? sr / n : 0.0;
}

void orc_power_iteration(int M, const int *row_ptr, const int *col_idx, const double *values,
                         double *x, double *y, int iters, double *lambdas) {
    for (int it = 0; it < iters; ++it) {
        memset(y, 0, (size_t)M * sizeof(double));
        orc_spmv_csr_serial(M, row_ptr, col_idx, values, x, y);
        double ss = 0.0;
        for (int r = 0; r < M; ++r) ss += y[r] * y[r];
        const double lam = sqrt(ss);
        lambdas[it] = lam;
        for (int r = 0; r < M; ++r) x[r] = y[r] / lam;
    }
}

/* ---- support for bench.py --impl reference at N > 1 (BASELINE config 5 at FULL size on the host) ---------------- */

/* The 7-point Laplacian on an n^3 grid, row r = (i n + j) n + k, columns r-n^2, r-n, r-1, r, r+1, r+n, r+n^2 where they
 * exist, values -1 x6 and 6 -- the same arrays as synth.lap3d_csr / the device generator
 * (tests/test_oracle_pinned.py compares them), written by all host threads: 938 M nonzeros in a few seconds instead of
 * minutes of numpy temporaries.  row_ptr[n^3 + 1]; col_idx / values sized by a first call with col_idx == NULL, which
 * only fills row_ptr and returns the number of nonzeros. */
long long orc_lap3d_csr(int n, int *row_ptr, int *col_idx, double *values) {
    const long long n2 = (long long)n * n, M = n2 * n;
    row_ptr[0] = 0;
    /* nnz of row (i, j, k) = 7 - faces touched; the prefix over a whole j-line is closed form, so every i-plane can
     * be numbered independently: rows before plane i hold 7 n^2 - 4 n (interior planes) or 6 n^2 - 4 n ... simpler and
     * still exact: count plane by plane (n numbers), then fill the planes in parallel. */
    long long *plane = malloc(((size_t)n + 1) * sizeof *plane);
    if (!plane) return -1;
    plane[0] = 0;
    for (int i = 0; i < n; ++i) {
        const long long per_plane = 7 * n2 - 2 * n /* j faces */ - 2 * n /* k faces */ - ((i == 0) + (i == n - 1)) * n2;
        plane[i + 1] = plane[i] + per_plane;
    }
    const long long nnz = plane[n];
    if (nnz > 0x7fffffffLL) {
        free(plane);
        return -1;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        long long at = plane[i];
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < n; ++k) {
                const long long r = ((long long)i * n + j) * n + k;
                const long long cand[7] = {r - n2, r - n, r - 1, r, r + 1, r + n, r + n2};
                const int on[7] = {i > 0, j > 0, k > 0, 1, k < n - 1, j < n - 1, i < n - 1};
                for (int e = 0; e < 7; ++e)
                    if (on[e]) {
                        if (col_idx) {
                            col_idx[at] = (int)cand[e];
                            values[at] = e == 3 ? 6.0 : -1.0;
                        }
                        ++at;
                    }
                row_ptr[r + 1] = (int)at;
            }
    }
    free(plane);
    (void)M;
    return nnz;
}

/* lambda = ||y||_2 ; x = y / lambda on all host threads (the norm and scale of one power iteration; the product is the
 * reference's own spvm_csr_parallel).  Fixed chunking, so the sum does not depend on the thread schedule. */
double orc_norm_scale(const double *y, double *x, long long n, int threads) {
    enum { CHUNKS = 1024 };
    double part[CHUNKS];
    const long long per = (n + CHUNKS - 1) / CHUNKS;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int c = 0; c < CHUNKS; ++c) {
        const long long lo = c * per, hi = lo + per < n ? lo + per : n;
        double s = 0.0;
        for (long long r = lo; r < hi; ++r) s += y[r] * y[r];
        part[c] = s;
    }
    double ss = 0.0;
    for (int c = 0; c < CHUNKS; ++c) ss += part[c];
    const double lam = sqrt(ss);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long long r = 0; r < n; ++r) x[r] = y[r] / lam;
    return lam;
}
