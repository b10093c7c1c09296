/*
 * matrix_parser.c -- drop-in read_matrix_market (reference src/matrix_parser.c:25-150).
 *
 * Same observable behaviour as the reference -- only sparse "matrix coordinate" files, 1-based to
 * 0-based, bounds check, pattern entries valued 1.0, and for 'S' (symmetric) files the mirrored
 * entry stored immediately after each off-diagonal entry; 0 on success, -1 with a message on
 * stdout otherwise -- but the body of the file is slurped once and tokenised in memory with
 * strtol/strtod (which accept exactly what fscanf's %d / %lf accept) instead of one fscanf call
 * per entry, and bodies of 4 MB or more are tokenised by all OpenMP threads at once.  This is row
 * (f).4 of SURVEY.md section 8: the fscanf loop is the end-to-end bottleneck on real files.
 */
#include "matrix_parser.h"

#include <ctype.h>
#include <errno.h>
#include <stdlib.h>
#include <string.h>

#include "utility.h"
#ifdef _OPENMP
#include <omp.h>
#endif

/* bodies at least this long are tokenised in parallel (SPMV_B200_PARSER_PARALLEL_MIN_BYTES overrides: tests) */
static size_t parallel_min_bytes(void) {
    const char *v = getenv("SPMV_B200_PARSER_PARALLEL_MIN_BYTES");
    if (v && *v) return (size_t)strtoull(v, NULL, 10);
    return (size_t)4 << 20;
}

void init_pre_matrix(PreMatrix *mat) {
    memset(mat, 0, sizeof *mat);
}

void free_pre_matrix(PreMatrix *mat) {
    FREE_CHECK(mat->I);
    FREE_CHECK(mat->J);
    FREE_CHECK(mat->val);
    mat->M = mat->N = mat->nz = 0;
}

static char *slurp_rest(FILE *f, size_t *len) {
    size_t cap = 1 << 16, n = 0;
    long here = ftell(f);
    if (here >= 0 && fseek(f, 0, SEEK_END) == 0) {
        long end = ftell(f);
        if (end >= here) cap = (size_t)(end - here) + 2;
        fseek(f, here, SEEK_SET);
    }
    char *buf = malloc(cap);
    if (!buf) return NULL;
    for (;;) {
        size_t got = fread(buf + n, 1, cap - 1 - n, f);
        n += got;
        if (got == 0) break;
        if (n == cap - 1) {
            char *bigger = realloc(buf, cap * 2);
            if (!bigger) { free(buf); return NULL; }
            buf = bigger;
            cap *= 2;
        }
    }
    buf[n] = '\0';
    *len = n;
    return buf;
}

/* %d semantics: skip white space, optional sign, at least one digit */
static int take_int(const char **cursor, int *out) {
    const char *p = *cursor;
    while (isspace((unsigned char)*p)) ++p;
    const char *q = p;
    if (*q == '+' || *q == '-') ++q;
    if (!isdigit((unsigned char)*q)) return 0;
    char *end;
    long v = strtol(p, &end, 10);
    *out = (int)v;
    *cursor = end;
    return 1;
}

/* %lf semantics */
static int take_double(const char **cursor, double *out) {
    const char *p = *cursor;
    while (isspace((unsigned char)*p)) ++p;
    if (*p == '\0') return 0;
    char *end;
    double v = strtod(p, &end);
    if (end == p) return 0;
    *out = v;
    *cursor = end;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Parallel tokenizer for large bodies.  Same result as the serial loop
 * below, which stays the reference for every irregular input: the parallel pass only accepts a
 * body made of at least `declared` x k white-space separated tokens (k = 2 for pattern files, else
 * 3) in which strtol / strtod consume every index / value token COMPLETELY and every index is in
 * bounds.  Anything else -- short file, "12abc", an index out of range -- returns 0 and the caller
 * re-parses serially, so messages and partial-read semantics are the reference's by construction.
 * Tokens, not lines, are the unit: fscanf("%d %d %lf") does not care about line breaks either.
 * ---------------------------------------------------------------------------------------- */
static int is_space(char ch) { return ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r' || ch == '\v' || ch == '\f'; }

static int parse_body_parallel(const char *text, size_t len, int declared, int pattern, int M, int N, int *I,
                               int *J, double *V) {
    const int k = pattern ? 2 : 3;
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    if (threads > 64) threads = 64;
    if (threads < 2 || declared <= 0) return 0;
    if (len < 2 * (size_t)threads) return 0; /* a tiny body: chunks would be empty (and chunk starts would sit at 0) */
    size_t begin[65];
    long long first_token[65];
    /* chunk boundaries moved forward to the start of a token (or the end of the text) */
    for (int t = 0; t <= threads; ++t) {
        size_t at = t == threads ? len : len / (size_t)threads * (size_t)t;
        if (t > 0 && t < threads) {
            while (at > 0 && at < len && !is_space(text[at - 1])) ++at; /* do not split a token */
        }
        begin[t] = at;
    }
    long long counts[64];
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; ++t) {
        long long n = 0;
        int in_token = 0;
        for (size_t p = begin[t]; p < begin[t + 1]; ++p) {
            const int sp = is_space(text[p]);
            if (!sp && !in_token) ++n;
            in_token = !sp;
        }
        counts[t] = n;
    }
    first_token[0] = 0;
    for (int t = 0; t < threads; ++t) first_token[t + 1] = first_token[t] + counts[t];
    const long long needed = (long long)declared * k;
    if (first_token[threads] < needed) return 0; /* short body: the serial loop reports the entry */
    int bad = 0;
#pragma omp parallel for num_threads(threads) schedule(static, 1) reduction(| : bad)
    for (int t = 0; t < threads; ++t) {
        long long tok = first_token[t];
        size_t p = begin[t];
        const size_t stop = begin[t + 1];
        while (p < stop && tok < needed && !bad) {
            while (p < stop && is_space(text[p])) ++p;
            if (p >= stop) break;
            size_t q = p;
            while (q < stop && !is_space(text[q])) ++q;
            const long long entry = tok / k;
            const int field = (int)(tok % k);
            char *end;
            if (field < 2) {
                const char *d = text + p;
                if (*d == '+' || *d == '-') ++d;
                if (*d < '0' || *d > '9') { bad = 1; break; }
                errno = 0;
                const long v = strtol(text + p, &end, 10);
                if (end != text + q || errno == ERANGE || v < 1 || v > (field == 0 ? M : N)) { bad = 1; break; }
                (field == 0 ? I : J)[entry] = (int)v - 1;
            } else {
                const double v = strtod(text + p, &end);
                if (end != text + q) { bad = 1; break; }
                V[entry] = v;
            }
            ++tok;
            p = q;
        }
    }
    if (bad) return 0;
    if (pattern) {
#pragma omp parallel for num_threads(threads)
        for (int e = 0; e < declared; ++e) V[e] = 1.0;
    }
    return 1;
}

/* in-place expansion of a symmetric body: entry e moves to e + (#off-diagonal entries before e), its
 * mirror right behind it (reference src/matrix_parser.c:100-118 stores the pair in this order) */
static size_t mirror_in_place(int declared, int *I, int *J, double *V) {
    size_t extra = 0;
    for (int e = 0; e < declared; ++e) extra += I[e] != J[e];
    size_t out = (size_t)declared + extra;
    for (int e = declared - 1; e >= 0; --e) { /* back to front: never overwrites an unread entry */
        const int r = I[e], c = J[e];
        const double v = V[e];
        if (r != c) {
            --out;
            I[out] = c; J[out] = r; V[out] = v;
        }
        --out;
        I[out] = r; J[out] = c; V[out] = v;
    }
    return (size_t)declared + extra;
}

int read_matrix_market(const char *filename, PreMatrix *mat) {
    FILE *f = fopen(filename, "r");
    if (!f) {
        printf("read_matrix_market: cannot open '%s'\n", filename);
        return -1;
    }
    if (mm_read_banner(f, &mat->type) != 0) {
        printf("read_matrix_market: '%s' has no valid Matrix Market banner\n", filename);
        fclose(f);
        return -1;
    }
    if (!mm_is_matrix(mat->type) || !mm_is_sparse(mat->type)) {
        printf("read_matrix_market: only sparse (coordinate) matrices are supported\n");
        fclose(f);
        return -1;
    }
    int declared = 0;
    if (mm_read_mtx_crd_size(f, &mat->M, &mat->N, &declared) != 0) {
        fclose(f);
        return -1;
    }
    size_t len = 0;
    char *text = slurp_rest(f, &len);
    fclose(f);
    if (!text) {
        printf("read_matrix_market: out of memory\n");
        return -1;
    }

    const int mirror = mm_is_symmetric(mat->type) ? 1 : 0;
    const int pattern = mm_is_pattern(mat->type) ? 1 : 0;
    const size_t cap = (size_t)(declared > 0 ? declared : 0) * (mirror ? 2u : 1u);
    int *I = malloc((cap ? cap : 1) * sizeof(int));
    int *J = malloc((cap ? cap : 1) * sizeof(int));
    double *V = malloc((cap ? cap : 1) * sizeof(double));
    int status = 0;
    size_t n = 0;
    if (!I || !J || !V) {
        printf("read_matrix_market: out of memory\n");
        status = -1;
    }
    const char *cur = text;
    int done = 0;
    if (status == 0 && len >= parallel_min_bytes()) {
        done = parse_body_parallel(text, len, declared, pattern, mat->M, mat->N, I, J, V);
        if (done) n = mirror ? mirror_in_place(declared, I, J, V) : (size_t)declared;
    }
    for (int e = 0; !done && status == 0 && e < declared; ++e) {
        int r, c;
        double v = 1.0;
        int fields = take_int(&cur, &r);
        if (fields == 1) fields += take_int(&cur, &c);
        if (fields == 2 && !pattern) fields += take_double(&cur, &v);
        if (fields != (pattern ? 2 : 3)) {
            printf("read_matrix_market: entry %d: read %d field(s) instead of %d\n", e + 1, fields,
                   pattern ? 2 : 3);
            status = -1;
            break;
        }
        --r;
        --c;
        if (r < 0 || r >= mat->M || c < 0 || c >= mat->N) {
            printf("read_matrix_market: index (%d,%d) outside a %dx%d matrix\n", r + 1, c + 1, mat->M, mat->N);
            status = -1;
            break;
        }
        I[n] = r; J[n] = c; V[n] = v; ++n;
        if (mirror && r != c) { I[n] = c; J[n] = r; V[n] = v; ++n; }
    }
    free(text);
    if (status != 0) {
        free(I); free(J); free(V);
        return -1;
    }
    /* hand out exact-size arrays, like the reference's final copy */
    mat->nz = (int)n;
    mat->I = malloc((n ? n : 1) * sizeof(int));
    mat->J = malloc((n ? n : 1) * sizeof(int));
    mat->val = malloc((n ? n : 1) * sizeof(double));
    if (!mat->I || !mat->J || !mat->val) {
        printf("read_matrix_market: out of memory\n");
        free(I); free(J); free(V);
        free_pre_matrix(mat);
        return -1;
    }
    memcpy(mat->I, I, n * sizeof(int));
    memcpy(mat->J, J, n * sizeof(int));
    memcpy(mat->val, V, n * sizeof(double));
    free(I); free(J); free(V);
    return 0;
}

void print_pre_matrix(PreMatrix *mat, const bool full_print) {
    char *kind = mm_typecode_to_str(mat->type);
    printf("matrix %d x %d, %d nonzeros, type: %s\n", mat->M, mat->N, mat->nz, kind ? kind : "?");
    free(kind);
    if (!full_print || mat->M > 30) return;
    printf("I  :"); for (int k = 0; k < mat->nz; ++k) printf(" %d", mat->I[k]); printf("\n");
    printf("J  :"); for (int k = 0; k < mat->nz; ++k) printf(" %d", mat->J[k]); printf("\n");
    printf("val:"); for (int k = 0; k < mat->nz; ++k) printf(" %f", mat->val[k]); printf("\n");
}
