/*
 * performance_calculate.h -- drop-in for reference libs/performance_calculate.h:10-67 (timing
 * accumulators, FLOPS = 2 nz / t, difference metrics), extended with the roofline figures the
 * B200 build reports (bytes moved, GB/s, fraction of the HBM peak).
 * The metric ids of the OpenMP build come first and keep their values; the ids of the CUDA
 * build (reference cuda_libs/performance_calculate.cuh:19-29) and the new kernels follow.
 */
#ifndef SPMV_B200_PERFORMANCE_CALCULATE_H
#define SPMV_B200_PERFORMANCE_CALCULATE_H
#include <stdbool.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define INITIAL_CAPACITY 100

typedef struct {
    double sum;
    double min;
    double max;
    double *values;
    double relative_error;
    double absolute_error;
    int count;
    int capacity;
} MetricStats;

typedef enum {
    SERIAL_TIME,
    PARALLEL_CSR_TIME,
    PARALLEL_SIMD_CSR_TIME,
    PARALLEL_HLL_TIME,
    PARALLEL_HLL_SIMD_TIME,
    SERIAL_HLL_TIME,
    /* reference CUDA build */
    ROW_CSR_TIME,
    WARP_CSR_TIME,
    ROW_HLL_TIME,
    WARP_HLL_TIME,
    WARP_SHARED_MEMORY_CSR_TIME,
    WARP_SHARED_MEMORY_HLL_TIME,
    /* this build */
    B200_CSR_TIME,
    B200_HLL_TIME,
    B200_CSR_E2E_TIME,
    B200_HLL_E2E_TIME,
    NUM_METRICS,
} MediumPerformanceMetric;

typedef struct DifferenceMetrics {
    double mean_abs_err;
    double mean_rel_err;
    int significant_diffs;
} DiffMetrics;

typedef struct performance_metrics {
    double time;
    double flops;
    double speedup;
    double efficiency;
} PerformanceMetrics;

struct DifferenceMetrics computeDifferenceMetrics(const double *ref, const double *res, int n, double abs_tol,
                                                  double rel_tol, bool print_summary);
void initialize_metrics(void);
void cleanup_metrics(void);
double get_metric_value(MediumPerformanceMetric type);
double get_relative_error(const MediumPerformanceMetric type);
double get_absolute_error(const MediumPerformanceMetric type);
void update_medium_metric(MediumPerformanceMetric type, double value);
void reset_medium_time_metrics(void);
DiffMetrics computeAverageErrors(const MediumPerformanceMetric type);
void accumulateErrors(const DiffMetrics *iteration_metrics, const MediumPerformanceMetric type);
double calculate_flops(int nz, double time);
void print_flops(double flops);

/* ---- additions --------------------------------------------------------------------------- */
double get_metric_min(MediumPerformanceMetric type);
double get_metric_max(MediumPerformanceMetric type);
double get_metric_median(MediumPerformanceMetric type);
/* algorithmic bytes of one product; value_bytes = 8 for fp64 */
long long calculate_csr_bytes(int M, int N, long long nnz, int value_bytes);
long long calculate_hll_bytes(int M, int N, long long slots, int num_blocks, int value_bytes);
double calculate_bandwidth_gbs(long long bytes, double time);
double calculate_roofline_fraction(long long bytes, double time, double peak_gbs);

#ifdef __cplusplus
}
#endif
#endif
