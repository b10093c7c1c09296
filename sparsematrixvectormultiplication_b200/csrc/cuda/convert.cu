// convert.cu -- format construction on the device: COO -> CSR.
//
// The host converter (csrc/host/csr_matrix.c, bit-exact with reference src/csr_matrix.c:63-126) costs a counting
// scatter plus one sort per row on one core; for matrices that are generated or already resident on the GPU the
// same CSR is built here with one stable radix sort of 64-bit (row, column) keys:
//     key = row << 32 | column,  payload = value      ->  col_idx = low word of the sorted keys, values = payload
//     row_ptr[r] = first position whose key is >= r << 32 (one binary search per row)
// Rows come out sorted by column.  For matrices without repeated coordinates the arrays equal convert_in_csr's bit
// for bit (a sorted duplicate-free row is unique).  Repeated coordinates keep their input order here, while the
// reference's order inside such a row is whatever its quicksort leaves (src/utility.c:38-91): same multiset, same
// product up to the rounding of a different summation order.  The CSR -> HLL step is spmv_b200_hll_from_csr (hll.cu).
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"
#include "handles.cuh"

extern "C" int spmv_b200_csr_adopt_device_(int M, int N, long long nnz, int *d_row_ptr, int *d_col_idx, double *d_values,
                                           void *stream, spmv_b200_csr **out);

namespace spmv {

__global__ void coo_keys_kernel(long long nz, int M, int N, const int *__restrict__ I, const int *__restrict__ J,
                                unsigned long long *__restrict__ keys, int *__restrict__ bad) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nz) return;
    const int r = I[k], c = J[k];
    if (r < 0 || r >= M || c < 0 || c >= N) atomicExch(bad, 1);
    keys[k] = ((unsigned long long)(unsigned int)r << 32) | (unsigned int)c;
}

__global__ void coo_finish_kernel(long long nz, int M, const unsigned long long *__restrict__ sorted, int *__restrict__ col_idx,
                                  int *__restrict__ row_ptr) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nz) col_idx[t] = (int)(sorted[t] & 0xffffffffULL);
    if (t <= M) {  // row_ptr[t] = number of keys below t << 32
        const unsigned long long bound = (unsigned long long)t << 32;
        long long lo = 0, hi = nz;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (sorted[mid] < bound) lo = mid + 1; else hi = mid;
        }
        row_ptr[t] = (int)lo;
    }
}

static int bits_for(int n) {
    int b = 1;
    while (b < 31 && (1LL << b) < n) ++b;
    return b;
}

}  // namespace spmv

using namespace spmv;

extern "C" {

int spmv_b200_csr_from_coo_device(int M, int N, long long nz, const int *d_I, const int *d_J, const double *d_val, void *stream_,
                                  spmv_b200_csr **out) {
    if (!out) return fail(SPMV_B200_ERR_INVALID, "csr_from_coo: out is NULL");
    *out = nullptr;
    if (M < 0 || N < 0 || nz < 0 || nz > 0x7fffffffLL || (nz > 0 && (!d_I || !d_J || !d_val)))
        return fail(SPMV_B200_ERR_INVALID, "csr_from_coo: bad arguments (M=%d N=%d nz=%lld)", M, N, nz);
    cudaStream_t stream = as_stream(stream_);
    unsigned long long *keys = nullptr, *sorted = nullptr;
    int *row_ptr = nullptr, *col_idx = nullptr, *d_bad = nullptr;
    double *values = nullptr;
    void *temp = nullptr;
    const size_t padded = std::max<size_t>(((size_t)nz + 3) & ~(size_t)3, 4);
    auto body = [&]() -> int {
        SPMV_TRY_CUDA(cudaMalloc(&row_ptr, ((size_t)M + 1) * sizeof(int)));
        SPMV_TRY_CUDA(cudaMalloc(&col_idx, padded * sizeof(int)));
        SPMV_TRY_CUDA(cudaMalloc(&values, padded * sizeof(double)));
        SPMV_TRY_CUDA(cudaMalloc(&keys, std::max<size_t>(nz, 1) * sizeof(unsigned long long)));
        SPMV_TRY_CUDA(cudaMalloc(&sorted, std::max<size_t>(nz, 1) * sizeof(unsigned long long)));
        SPMV_TRY_CUDA(cudaMalloc(&d_bad, sizeof(int)));
        SPMV_TRY_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), stream));
        SPMV_TRY_CUDA(cudaMemsetAsync(col_idx, 0, padded * sizeof(int), stream));
        SPMV_TRY_CUDA(cudaMemsetAsync(values, 0, padded * sizeof(double), stream));
        if (nz > 0) {
            coo_keys_kernel<<<blocks_for(nz, 256), 256, 0, stream>>>(nz, M, N, d_I, d_J, keys, d_bad);
            SPMV_TRY_CUDA(cudaGetLastError());
            const int end_bit = 32 + bits_for(std::max(M, 1));
            size_t temp_bytes = 0;
            SPMV_TRY_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys, sorted, d_val, values, (int)nz, 0, end_bit, stream));
            SPMV_TRY_CUDA(cudaMalloc(&temp, temp_bytes ? temp_bytes : 1));
            SPMV_TRY_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, sorted, d_val, values, (int)nz, 0, end_bit, stream));
        }
        coo_finish_kernel<<<blocks_for(std::max<long long>(nz, (long long)M + 1), 256), 256, 0, stream>>>(nz, M, sorted, col_idx, row_ptr);
        SPMV_TRY_CUDA(cudaGetLastError());
        int bad = 0;
        SPMV_TRY_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
        SPMV_TRY_CUDA(cudaStreamSynchronize(stream));
        if (bad) return fail(SPMV_B200_ERR_INVALID, "csr_from_coo: an index lies outside the %d x %d matrix", M, N);
        return SPMV_B200_OK;
    };
    int rc = body();
    cudaFree(keys);
    cudaFree(sorted);
    cudaFree(d_bad);
    cudaFree(temp);
    if (rc == SPMV_B200_OK) rc = spmv_b200_csr_adopt_device_(M, N, nz, row_ptr, col_idx, values, stream_, out);
    if (rc != SPMV_B200_OK) {
        cudaFree(row_ptr);
        cudaFree(col_idx);
        cudaFree(values);
    }
    return rc;
}

int spmv_b200_csr_from_coo(int M, int N, long long nz, const int *I, const int *J, const double *val, spmv_b200_csr **out) {
    if (!out) return fail(SPMV_B200_ERR_INVALID, "csr_from_coo: out is NULL");
    *out = nullptr;
    if (nz < 0 || (nz > 0 && (!I || !J || !val))) return fail(SPMV_B200_ERR_INVALID, "csr_from_coo: bad arguments");
    int *d_I = nullptr, *d_J = nullptr;
    double *d_val = nullptr;
    auto body = [&]() -> int {
        const size_t n = std::max<size_t>(nz, 1);
        SPMV_TRY_CUDA(cudaMalloc(&d_I, n * sizeof(int)));
        SPMV_TRY_CUDA(cudaMalloc(&d_J, n * sizeof(int)));
        SPMV_TRY_CUDA(cudaMalloc(&d_val, n * sizeof(double)));
        if (nz > 0) {
            SPMV_TRY_CUDA(cudaMemcpy(d_I, I, (size_t)nz * sizeof(int), cudaMemcpyHostToDevice));
            SPMV_TRY_CUDA(cudaMemcpy(d_J, J, (size_t)nz * sizeof(int), cudaMemcpyHostToDevice));
            SPMV_TRY_CUDA(cudaMemcpy(d_val, val, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice));
        }
        return spmv_b200_csr_from_coo_device(M, N, nz, d_I, d_J, d_val, nullptr, out);
    };
    const int rc = body();
    cudaFree(d_I);
    cudaFree(d_J);
    cudaFree(d_val);
    return rc;
}

}  // extern "C"
