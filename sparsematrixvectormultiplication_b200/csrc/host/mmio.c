/*
 * mmio.c -- Matrix Market banner / size-line reader, API compatible with the NIST mmio subset
 * used by the reference (reference libs/mmio.c:96-214; semantics summarised in SURVEY.md
 * section 2 row 2).  Independent implementation: one tokenising pass over the banner line and a
 * table lookup per field instead of the chain of strcmp branches.
 */
#include "mmio.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>

struct field_name { const char *word; char code; };

static const struct field_name k_format[] = {{"coordinate", 'C'}, {"array", 'A'}, {NULL, 0}};
static const struct field_name k_data[] = {{"real", 'R'}, {"complex", 'C'}, {"pattern", 'P'}, {"integer", 'I'}, {NULL, 0}};
static const struct field_name k_symmetry[] = {{"general", 'G'}, {"symmetric", 'S'}, {"hermitian", 'H'},
                                               {"skew-symmetric", 'K'}, {NULL, 0}};

static char lookup(const struct field_name *table, const char *word) {
    for (; table->word; ++table)
        if (strcmp(table->word, word) == 0) return table->code;
    return 0;
}

static const char *reverse_lookup(const struct field_name *table, char code) {
    for (; table->word; ++table)
        if (table->code == code) return table->word;
    return NULL;
}

/* next whitespace-delimited token of at most MM_MAX_TOKEN_LENGTH-1 chars, lower-cased on demand */
static const char *next_token(const char *p, char *out, int fold) {
    while (*p && isspace((unsigned char)*p)) ++p;
    int n = 0;
    while (*p && !isspace((unsigned char)*p)) {
        if (n < MM_MAX_TOKEN_LENGTH - 1) out[n++] = fold ? (char)tolower((unsigned char)*p) : *p;
        ++p;
    }
    out[n] = '\0';
    return n ? p : NULL;
}

int mm_read_banner(FILE *f, MM_typecode *matcode) {
    char line[MM_MAX_LINE_LENGTH];
    char tok[5][MM_MAX_TOKEN_LENGTH];
    mm_clear_typecode(matcode);
    if (fgets(line, sizeof line, f) == NULL) return MM_PREMATURE_EOF;
    const char *p = line;
    for (int k = 0; k < 5; ++k) {
        p = next_token(p, tok[k], k > 0);
        if (!p) return MM_PREMATURE_EOF; /* fewer than five fields on the banner line */
    }
    if (strncmp(tok[0], MatrixMarketBanner, strlen(MatrixMarketBanner)) != 0) return MM_NO_HEADER;
    if (strcmp(tok[1], "matrix") != 0) return MM_UNSUPPORTED_TYPE;
    (*matcode)[0] = 'M';
    const char fmt = lookup(k_format, tok[2]);
    if (!fmt) return MM_UNSUPPORTED_TYPE;
    (*matcode)[1] = fmt;
    const char data = lookup(k_data, tok[3]);
    if (!data) return MM_UNSUPPORTED_TYPE;
    (*matcode)[2] = data;
    const char sym = lookup(k_symmetry, tok[4]);
    if (!sym) return MM_UNSUPPORTED_TYPE;
    (*matcode)[3] = sym;
    return 0;
}

int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz) {
    char line[MM_MAX_LINE_LENGTH];
    *M = *N = *nz = 0;
    do { /* comments start with '%' in column one */
        if (fgets(line, sizeof line, f) == NULL) return MM_PREMATURE_EOF;
    } while (line[0] == '%');
    if (sscanf(line, "%d %d %d", M, N, nz) == 3) return 0;
    /* the first non-comment line was blank: the three integers follow somewhere below.
     * (The reference loops forever when a non-numeric token shows up here; we report EOF.) */
    int got = fscanf(f, "%d %d %d", M, N, nz);
    return got == 3 ? 0 : MM_PREMATURE_EOF;
}

int mm_is_valid(MM_typecode t) {
    if (!mm_is_matrix(t)) return 0;
    if (mm_is_dense(t) && mm_is_pattern(t)) return 0;
    if (mm_is_real(t) && mm_is_hermitian(t)) return 0;
    if (mm_is_pattern(t) && (mm_is_hermitian(t) || mm_is_skew(t))) return 0;
    return 1;
}

char *mm_typecode_to_str(MM_typecode t) {
    const char *obj = mm_is_matrix(t) ? "matrix" : NULL;
    const char *fmt = reverse_lookup(k_format, t[1]);
    const char *data = reverse_lookup(k_data, t[2]);
    const char *sym = reverse_lookup(k_symmetry, t[3]);
    if (!obj || !fmt || !data || !sym) return NULL;
    char *out = malloc(MM_MAX_LINE_LENGTH);
    if (out) snprintf(out, MM_MAX_LINE_LENGTH, "%s %s %s %s", obj, fmt, data, sym);
    return out;
}

int mm_write_banner(FILE *f, MM_typecode t) {
    char *s = mm_typecode_to_str(t);
    if (!s) return MM_COULD_NOT_WRITE_FILE;
    int rc = fprintf(f, "%s %s\n", MatrixMarketBanner, s);
    free(s);
    return rc < 0 ? MM_COULD_NOT_WRITE_FILE : 0;
}

int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz) {
    return fprintf(f, "%d %d %d\n", M, N, nz) < 0 ? MM_COULD_NOT_WRITE_FILE : 0;
}
