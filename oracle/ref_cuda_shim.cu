// ref_cuda_shim.cu -- TEST / BENCH INFRASTRUCTURE (never linked into the product).
//
// Launches the reference's OWN GPU kernels -- compiled, unmodified, where they lie under /root/reference by
// oracle/Makefile (target ref_cuda) -- on device arrays handed in through a plain C interface, with the launch shapes
// the reference driver uses (main_cuda.cu:148-151, :212-222, :285-306, :545-550, :613-623).  It is the "existing
// kernel recompiled for sm_100a" bar that bench.py reports next to the new kernels, and a second checker for the GPU
// parity tests.  Nothing here is the reference's code: the kernels are declared by the reference's headers
// (cuda_libs/*.cuh, included from their own directory) and linked from its objects.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "csr_matrix_cuda.cuh"  // reference cuda_libs/: kernel prototypes, CSRMatrix
#include "hll_matrix.cuh"       // reference cuda_libs/: kernel prototypes, ELLPACKBlock, HLLMatrix

#define SHIM_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            std::fprintf(stderr, "ref_cuda_shim: %s: %s\n", #expr, cudaGetErrorString(e__)); \
            return -1;                                                                     \
        }                                                                                  \
    } while (0)

struct RefHll {  // device image in the REFERENCE's layout: array of blocks with row-major JA / AS
    int M = 0, num_blocks = 0, max_maxnz = 0;
    ELLPACKBlock *d_blocks = nullptr;
    int *arena_ja = nullptr;
    double *arena_as = nullptr;
};

extern "C" {

// which: 0 = spmv_csr_naive_kernel, 1 = spmv_csr_warp_kernel, 2 = spmv_csr_warp_shared_memory_kernel
int ref_cuda_csr_spmv(int which, int M, int N, const int *d_row_ptr, const int *d_col_idx, const double *d_values,
                      const double *d_x, double *d_y, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int minGrid = 0, blockSize = 0;
    if (M <= 0) return 0;
    if (which == 0) {
        SHIM_TRY(cudaOccupancyMaxPotentialBlockSize(&minGrid, &blockSize, spmv_csr_naive_kernel, 0, M));
        spmv_csr_naive_kernel<<<(M + blockSize - 1) / blockSize, blockSize, 0, stream>>>(M, d_row_ptr, d_col_idx, d_values, d_x, d_y);
    } else if (which == 1) {
        SHIM_TRY(cudaOccupancyMaxPotentialBlockSize(&minGrid, &blockSize, spmv_csr_warp_kernel, 0, 0));
        const int warps = blockSize / 32;
        spmv_csr_warp_kernel<<<(M + warps - 1) / warps, blockSize, 0, stream>>>(M, d_row_ptr, d_col_idx, d_values, d_x, d_y);
    } else if (which == 2) {
        const int cache = std::min(N, MAX_CACHE);
        const size_t smem = sizeof(double) * cache;
        SHIM_TRY(cudaOccupancyMaxPotentialBlockSizeVariableSMem(&minGrid, &blockSize, spmv_csr_warp_shared_memory_kernel,
                                                                [smem](int) { return smem; }, 0));
        const int warps = blockSize / 32;
        spmv_csr_warp_shared_memory_kernel<<<(M + warps - 1) / warps, blockSize, smem, stream>>>(M, d_row_ptr, d_col_idx, d_values,
                                                                                              d_x, d_y, cache);
    } else {
        return -1;
    }
    SHIM_TRY(cudaGetLastError());
    return 0;
}

// Builds the reference's device layout from a host HLLMatrix (reference layout).  One arena instead of the
// reference driver's two cudaMalloc per block (main_cuda.cu:369-402): the kernels accept it unchanged.
int ref_cuda_hll_upload(const HLLMatrix *hll, int M, RefHll **out) {
    RefHll *H = new RefHll();
    H->M = M;
    H->num_blocks = hll->num_blocks;
    std::vector<ELLPACKBlock> blocks(hll->num_blocks);
    size_t total = 0;
    for (int b = 0; b < hll->num_blocks; ++b) total += (size_t)hll->blocks[b].M * hll->blocks[b].MAXNZ;
    SHIM_TRY(cudaMalloc(&H->arena_ja, std::max<size_t>(total, 1) * sizeof(int)));
    SHIM_TRY(cudaMalloc(&H->arena_as, std::max<size_t>(total, 1) * sizeof(double)));
    std::vector<int> ja(std::max<size_t>(total, 1));
    std::vector<double> as(std::max<size_t>(total, 1));
    size_t at = 0;
    for (int b = 0; b < hll->num_blocks; ++b) {
        const ELLPACKBlock &src = hll->blocks[b];
        const size_t n = (size_t)src.M * src.MAXNZ;
        blocks[b] = src;
        blocks[b].JA = H->arena_ja + at;
        blocks[b].AS = H->arena_as + at;
        if (n) {
            std::copy(src.JA, src.JA + n, ja.begin() + at);
            std::copy(src.AS, src.AS + n, as.begin() + at);
        }
        H->max_maxnz = std::max(H->max_maxnz, src.MAXNZ);
        at += n;
    }
    SHIM_TRY(cudaMemcpy(H->arena_ja, ja.data(), total * sizeof(int), cudaMemcpyHostToDevice));
    SHIM_TRY(cudaMemcpy(H->arena_as, as.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    SHIM_TRY(cudaMalloc(&H->d_blocks, std::max<size_t>(blocks.size(), 1) * sizeof(ELLPACKBlock)));
    SHIM_TRY(cudaMemcpy(H->d_blocks, blocks.data(), blocks.size() * sizeof(ELLPACKBlock), cudaMemcpyHostToDevice));
    *out = H;
    return 0;
}

// which: 0 = spmv_hll_naive_kernel, 1 = spmv_hll_warp_kernel, 2 = spmv_hll_warp_shared_kernel_v1
int ref_cuda_hll_spmv(const RefHll *H, int which, const double *d_x, double *d_y, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int minGrid = 0, blockSize = 0;
    const int M = H->M;
    if (M <= 0) return 0;
    if (which == 0) {
        SHIM_TRY(cudaOccupancyMaxPotentialBlockSize(&minGrid, &blockSize, spmv_hll_naive_kernel, 0, 0));
        spmv_hll_naive_kernel<<<(M + blockSize - 1) / blockSize, blockSize, 0, stream>>>(M, H->d_blocks, d_x, d_y);
    } else if (which == 1) {
        SHIM_TRY(cudaOccupancyMaxPotentialBlockSize(&minGrid, &blockSize, spmv_hll_warp_kernel, 0, 0));
        const int warps = blockSize / 32;
        spmv_hll_warp_kernel<<<(M + warps - 1) / warps, blockSize, 0, stream>>>(M, H->d_blocks, d_x, d_y);
    } else if (which == 2) {
        const int max_w = H->max_maxnz;
        SHIM_TRY(cudaOccupancyMaxPotentialBlockSizeVariableSMem(&minGrid, &blockSize, spmv_hll_warp_shared_kernel_v1,
                                                                [max_w](int b) { return (size_t)(b / 32) * max_w * sizeof(double); }, 0));
        const int warps = blockSize / 32;
        const size_t smem = (size_t)warps * max_w * sizeof(double);
        spmv_hll_warp_shared_kernel_v1<<<(M + warps - 1) / warps, blockSize, smem, stream>>>(M, H->d_blocks, d_x, d_y);
    } else {
        return -1;
    }
    SHIM_TRY(cudaGetLastError());
    return 0;
}

void ref_cuda_hll_free(RefHll *H) {
    if (!H) return;
    cudaFree(H->d_blocks);
    cudaFree(H->arena_ja);
    cudaFree(H->arena_as);
    delete H;
}

}  // extern "C"
