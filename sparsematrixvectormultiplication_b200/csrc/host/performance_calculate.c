/*
 * performance_calculate.c -- timing accumulators, FLOPS and difference metrics with the
 * signatures of reference src/performance_calculate.c:11-178, plus the roofline arithmetic the
 * B200 build reports.
 */
#include "performance_calculate.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "utility.h"

static MetricStats g_metrics[NUM_METRICS]; /* process-global, like the reference's metrics[] */

static int valid(MediumPerformanceMetric t) { return (int)t >= 0 && t < NUM_METRICS; }

void initialize_metrics(void) {
    for (int m = 0; m < NUM_METRICS; ++m) {
        free(g_metrics[m].values);
        memset(&g_metrics[m], 0, sizeof g_metrics[m]);
        g_metrics[m].capacity = INITIAL_CAPACITY;
        g_metrics[m].values = malloc(INITIAL_CAPACITY * sizeof(double));
        if (!g_metrics[m].values) g_metrics[m].capacity = 0;
    }
}

void cleanup_metrics(void) {
    for (int m = 0; m < NUM_METRICS; ++m) {
        free(g_metrics[m].values);
        g_metrics[m].values = NULL;
        g_metrics[m].capacity = 0;
    }
}

void reset_medium_time_metrics(void) {
    for (int m = 0; m < NUM_METRICS; ++m) {
        g_metrics[m].sum = 0.0;
        g_metrics[m].count = 0;
        g_metrics[m].relative_error = g_metrics[m].absolute_error = 0.0;
    }
}

void update_medium_metric(MediumPerformanceMetric type, double value) {
    if (!valid(type)) return;
    MetricStats *s = &g_metrics[type];
    if (s->count == 0 || value < s->min) s->min = value;
    if (s->count == 0 || value > s->max) s->max = value;
    s->sum += value;
    if (s->count >= s->capacity) {
        int grown = s->capacity > 0 ? 2 * s->capacity : INITIAL_CAPACITY;
        double *v = realloc(s->values, (size_t)grown * sizeof(double));
        if (v) { s->values = v; s->capacity = grown; }
    }
    if (s->count < s->capacity) s->values[s->count] = value;
    s->count++;
}

double get_metric_value(MediumPerformanceMetric type) { /* mean of the recorded samples */
    if (!valid(type) || g_metrics[type].count == 0) return 0.0;
    return g_metrics[type].sum / g_metrics[type].count;
}

double get_metric_min(MediumPerformanceMetric type) {
    return valid(type) && g_metrics[type].count ? g_metrics[type].min : 0.0;
}

double get_metric_max(MediumPerformanceMetric type) {
    return valid(type) && g_metrics[type].count ? g_metrics[type].max : 0.0;
}

static int cmp_double(const void *a, const void *b) {
    const double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

double get_metric_median(MediumPerformanceMetric type) {
    if (!valid(type)) return 0.0;
    const MetricStats *s = &g_metrics[type];
    const int n = s->count < s->capacity ? s->count : s->capacity;
    if (n == 0 || !s->values) return 0.0;
    double *tmp = malloc((size_t)n * sizeof(double));
    if (!tmp) return 0.0;
    memcpy(tmp, s->values, (size_t)n * sizeof(double));
    qsort(tmp, (size_t)n, sizeof(double), cmp_double);
    const double med = (n & 1) ? tmp[n / 2] : 0.5 * (tmp[n / 2 - 1] + tmp[n / 2]);
    free(tmp);
    return med;
}

double get_relative_error(const MediumPerformanceMetric type) {
    return valid(type) && g_metrics[type].count ? g_metrics[type].relative_error : 0.0;
}

double get_absolute_error(const MediumPerformanceMetric type) {
    return valid(type) && g_metrics[type].count ? g_metrics[type].absolute_error : 0.0;
}

void accumulateErrors(const DiffMetrics *iteration_metrics, const MediumPerformanceMetric type) {
    if (!valid(type)) return;
    g_metrics[type].absolute_error += iteration_metrics->mean_abs_err;
    g_metrics[type].relative_error += iteration_metrics->mean_rel_err;
}

/* errors are accumulated on every iteration, timings only after the ITERATION_SKIP warm-ups,
 * hence the "+ ITERATION_SKIP" in the divisor (reference src/performance_calculate.c:58-67) */
DiffMetrics computeAverageErrors(const MediumPerformanceMetric type) {
    DiffMetrics avg = {0.0, 0.0, 0};
    if (valid(type) && g_metrics[type].count > 0) {
        const double runs = g_metrics[type].count + ITERATION_SKIP;
        avg.mean_abs_err = g_metrics[type].absolute_error / runs;
        avg.mean_rel_err = g_metrics[type].relative_error / runs;
    }
    return avg;
}

double calculate_flops(int nz, double time) { return 2.0 * nz / time; }

void print_flops(double flops) {
    static const char *unit[] = {"FLOPS", "KFLOPS", "MFLOPS", "GFLOPS", "TFLOPS", "PFLOPS", "EFLOPS"};
    int u = 0;
    for (; flops >= 1000.0 && u < 6; ++u) flops /= 1000.0;
    printf("%.3f %s\n", flops, unit[u]);
}

/* C-build rule: an entry counts as a significant difference when |ref-res| > abs_tol AND the
 * relative difference (denominator max(|ref|,|res|,rel_tol)) exceeds rel_tol; the mean is
 * taken over those entries only and mean_abs_err stays 0. */
struct DifferenceMetrics computeDifferenceMetrics(const double *ref, const double *res, int n, double abs_tol,
                                                  double rel_tol, bool print_summary) {
    struct DifferenceMetrics out = {0.0, 0.0, 0};
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        const double gap = fabs(ref[i] - res[i]);
        if (!(gap > abs_tol)) continue;
        const double scale = fmax(fmax(fabs(ref[i]), fabs(res[i])), rel_tol);
        const double rel = gap / scale;
        if (rel > rel_tol) {
            acc += rel;
            out.significant_diffs++;
        }
    }
    if (out.significant_diffs > 0) out.mean_rel_err = acc / out.significant_diffs;
    if (print_summary) {
        printf("--- Comparison Summary ---\n");
        printf("Vector size : %d\n", n > 0 ? n : 0);
        if (n <= 0) printf("Result : PASS (empty vectors)\n");
        else {
            printf("Significant differences : %d\n", out.significant_diffs);
            printf("Mean Significant Relative Error : %.10e\n", out.mean_rel_err);
        }
        printf("----------------------------\n");
    }
    return out;
}

long long calculate_csr_bytes(int M, int N, long long nnz, int value_bytes) {
    return nnz * (value_bytes + 4) + 4LL * ((long long)M + 1) + (long long)value_bytes * M +
           (long long)value_bytes * N;
}

long long calculate_hll_bytes(int M, int N, long long slots, int num_blocks, int value_bytes) {
    return slots * (value_bytes + 4) + 8LL * ((long long)num_blocks + 1) + (long long)value_bytes * M +
           (long long)value_bytes * N;
}

double calculate_bandwidth_gbs(long long bytes, double time) { return (double)bytes / time / 1e9; }

double calculate_roofline_fraction(long long bytes, double time, double peak_gbs) {
    return calculate_bandwidth_gbs(bytes, time) / peak_gbs;
}
