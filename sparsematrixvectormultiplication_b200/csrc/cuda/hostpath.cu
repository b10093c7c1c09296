// hostpath.cu -- the end-to-end entry points spmv_b200_{csr,hll}_spmv_host: host x -> device, product,
// device -> host y, as ONE pipelined pass.
//
// They replace the reference driver's "cudaMemcpy(d_x) ... kernel ... cudaMemcpy(y, DeviceToHost)" sequence
// (reference main_cuda.cu:145,166,183 and :454,471).  A host-buffer product of a resident matrix is bound by the
// PCIe link, not by HBM: 8 bytes up per column and 8 bytes down per row against 12 bytes per nonzero streamed at
// ~100x the speed.  The link is full duplex, so the call is organised to keep BOTH directions busy:
//
//   * the rows are cut into W windows (whole tiles of the plan, i.e. equal bytes of matrix stream);
//   * x travels in W chunks on an upload stream; window w is launched on a compute stream as soon as the chunk that
//     holds its largest referenced column has landed (need[w] = running maximum of max(col)+1 over windows 0..w,
//     found once per matrix by a reduction over col_idx / JA);
//   * the rows of window w are copied back on a download stream as soon as its kernel is done.
//
// For banded matrices (stencils, FEM) the upload of chunk w+1, the product of window w and the download of window
// w-1 overlap, and the call takes max(upload, download) instead of their sum.  For matrices whose rows reference
// the whole of x (random columns) the windows all wait for the last chunk: the upload is serial, the downloads
// still overlap the products.  Results do not depend on W: a window launch runs the same kernel over a sub-range
// of the same tiles.  Pageable host buffers work but serialise (cudaMemcpyAsync stages them); pinned or
// cudaHostRegister'ed buffers give the overlap.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "handles.cuh"

namespace spmv {

constexpr int kMaxWindows = 64;  // events; the automatic choice stops at kAutoWindows.  Round 1, equal windows, measured on lap2d 4096^2: 4 -> 3.93 ms, 8 -> 3.49, 16 -> 3.42, 24 -> 3.91, 32 -> 3.65, 48 -> 3.89
constexpr int kAutoWindows = 16;
constexpr long long kMinWindowBytes = 2LL << 20;  // below 2 MB of x + y per window the launch overheads win

struct HostPipe {
    cudaStream_t up = nullptr, compute = nullptr, down = nullptr;
    std::vector<cudaEvent_t> landed;  // chunk c of x is on the device
    std::vector<cudaEvent_t> done;    // window w is computed
    cudaEvent_t y_landed = nullptr;   // accumulate: the old y is on the device
    // window plan (valid for `path` only)
    int path = -1;
    int windows = 0;
    std::vector<int> unit;        // W+1: first tile / row / hack of every window
    std::vector<long long> row;   // W+1: first row of every window
    std::vector<long long> need;  // W: doubles of x that must be resident before window w starts (running maximum)
};

void host_pipe_free(HostPipe *p) {
    if (!p) return;
    for (cudaEvent_t e : p->landed) cudaEventDestroy(e);
    for (cudaEvent_t e : p->done) cudaEventDestroy(e);
    if (p->y_landed) cudaEventDestroy(p->y_landed);
    if (p->up) cudaStreamDestroy(p->up);
    if (p->compute) cudaStreamDestroy(p->compute);
    if (p->down) cudaStreamDestroy(p->down);
    delete p;
}

static int pipe_create(HostPipe **out) {
    HostPipe *p = new (std::nothrow) HostPipe();
    if (!p) return fail(SPMV_B200_ERR_NOMEM, "spmv_host: out of host memory");
    auto body = [&]() -> int {
        SPMV_TRY_CUDA(cudaStreamCreateWithFlags(&p->up, cudaStreamNonBlocking));
        SPMV_TRY_CUDA(cudaStreamCreateWithFlags(&p->compute, cudaStreamNonBlocking));
        SPMV_TRY_CUDA(cudaStreamCreateWithFlags(&p->down, cudaStreamNonBlocking));
        SPMV_TRY_CUDA(cudaEventCreateWithFlags(&p->y_landed, cudaEventDisableTiming));
        p->landed.resize(kMaxWindows, nullptr);
        p->done.resize(kMaxWindows, nullptr);
        for (int i = 0; i < kMaxWindows; ++i) {
            SPMV_TRY_CUDA(cudaEventCreateWithFlags(&p->landed[i], cudaEventDisableTiming));
            SPMV_TRY_CUDA(cudaEventCreateWithFlags(&p->done[i], cudaEventDisableTiming));
        }
        return SPMV_B200_OK;
    };
    const int rc = body();
    if (rc != SPMV_B200_OK) {
        host_pipe_free(p);
        return rc;
    }
    *out = p;
    return SPMV_B200_OK;
}

// out[w] = max(idx[k]) + 1 over k in [bound[w], bound[w+1]); 0 for an empty range.  One launch for all windows.
__global__ void window_need_kernel(const int *__restrict__ idx, const long long *__restrict__ bound, int windows,
                                   int *__restrict__ out) {
    const int w = blockIdx.y;
    const long long lo = bound[w], hi = bound[w + 1];
    int top = 0;
    for (long long k = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += (long long)gridDim.x * blockDim.x)
        top = max(top, idx[k] + 1);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, off));
    if ((threadIdx.x & 31) == 0 && top > 0) atomicMax(out + w, top);  // integer max: order independent
}

// need[] from the element ranges of the windows (elements = nonzeros of CSR / slots of HLL)
static int plan_needs(HostPipe *p, const int *d_idx, const std::vector<long long> &elem_bound) {
    const int W = p->windows;
    p->need.assign(W, 0);
    if (W == 0 || elem_bound[W] == elem_bound[0]) return SPMV_B200_OK;
    long long *d_bound = nullptr;
    int *d_out = nullptr;
    std::vector<int> top(W, 0);
    auto body = [&]() -> int {
        SPMV_TRY_CUDA(cudaMalloc(&d_bound, (size_t)(W + 1) * sizeof(long long)));
        SPMV_TRY_CUDA(cudaMalloc(&d_out, (size_t)W * sizeof(int)));
        SPMV_TRY_CUDA(cudaMemcpyAsync(d_bound, elem_bound.data(), (size_t)(W + 1) * sizeof(long long), cudaMemcpyHostToDevice, p->compute));
        SPMV_TRY_CUDA(cudaMemsetAsync(d_out, 0, (size_t)W * sizeof(int), p->compute));
        window_need_kernel<<<dim3(296, W), 256, 0, p->compute>>>(d_idx, d_bound, W, d_out);
        SPMV_TRY_CUDA(cudaGetLastError());
        SPMV_TRY_CUDA(cudaMemcpyAsync(top.data(), d_out, (size_t)W * sizeof(int), cudaMemcpyDeviceToHost, p->compute));
        SPMV_TRY_CUDA(cudaStreamSynchronize(p->compute));
        return SPMV_B200_OK;
    };
    const int rc = body();
    cudaFree(d_bound);
    cudaFree(d_out);
    if (rc != SPMV_B200_OK) return rc;
    long long running = 0;
    for (int w = 0; w < W; ++w) {
        running = std::max<long long>(running, top[w]);
        p->need[w] = running;
    }
    return SPMV_B200_OK;
}

static int pick_windows(long long M, long long N, int units) {
    const int forced = env_int("SPMV_B200_HOST_WINDOWS", 0);
    long long w = forced > 0 ? forced : (8 * (M + N)) / kMinWindowBytes;
    w = std::max<long long>(1, std::min<long long>(w, forced > 0 ? kMaxWindows : kAutoWindows));
    return (int)std::max<long long>(1, std::min<long long>(w, units));
}

// The pass itself.  launch(w) enqueues the kernel(s) of window w on p->compute.
template <class Launch>
static int run_pipeline(HostPipe *p, long long M, long long N, const double *x, double *y, double *d_x, double *d_y,
                        bool accumulate, Launch launch) {
    const int W = p->windows;
    // (1) uploads: the old y first (accumulate), then x in W chunks
    if (accumulate && M) {
        SPMV_TRY_CUDA(cudaMemcpyAsync(d_y, y, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, p->up));
        SPMV_TRY_CUDA(cudaEventRecord(p->y_landed, p->up));
    }
    const long long chunk = W > 0 ? (N + W - 1) / W : N;
    int chunks = 0;
    for (long long at = 0; at < N; at += chunk, ++chunks) {
        const long long n = std::min(chunk, N - at);
        SPMV_TRY_CUDA(cudaMemcpyAsync(d_x + at, x + at, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, p->up));
        SPMV_TRY_CUDA(cudaEventRecord(p->landed[chunks], p->up));
    }
    // (2) products: window w waits for the chunk that completes its referenced prefix of x
    for (int w = 0; w < W; ++w) {
        if (accumulate && M && w == 0) SPMV_TRY_CUDA(cudaStreamWaitEvent(p->compute, p->y_landed, 0));
        if (chunks > 0 && p->need[w] > 0) {
            const int c = (int)std::min<long long>((p->need[w] - 1) / chunk, chunks - 1);
            SPMV_TRY_CUDA(cudaStreamWaitEvent(p->compute, p->landed[c], 0));
        }
        SPMV_TRY(launch(w));
        SPMV_TRY_CUDA(cudaEventRecord(p->done[w], p->compute));
    }
    // (3) downloads, window by window
    for (int w = 0; w < W; ++w) {
        const long long r0 = p->row[w], r1 = p->row[w + 1];
        if (r1 <= r0) continue;
        SPMV_TRY_CUDA(cudaStreamWaitEvent(p->down, p->done[w], 0));
        SPMV_TRY_CUDA(cudaMemcpyAsync(y + r0, d_y + r0, (size_t)(r1 - r0) * sizeof(double), cudaMemcpyDeviceToHost, p->down));
    }
    SPMV_TRY_CUDA(cudaStreamSynchronize(p->down));
    SPMV_TRY_CUDA(cudaStreamSynchronize(p->compute));
    SPMV_TRY_CUDA(cudaStreamSynchronize(p->up));
    return SPMV_B200_OK;
}

static void abort_pipeline(HostPipe *p) {  // after an error: nothing of this call may still be in flight
    if (!p) return;
    cudaStreamSynchronize(p->up);
    cudaStreamSynchronize(p->compute);
    cudaStreamSynchronize(p->down);
}

// ---- CSR --------------------------------------------------------------------------------------------------------
static int csr_plan_windows(spmv_b200_csr *A, CsrPath path) {
    HostPipe *p = A->pipe;
    const bool by_tiles = path != kPathVector && path != kPathRow;
    // long rows are computed after every tile (they need all of x and are written last), the binned kernel walks the
    // rows in bin order: one window
    const int units = by_tiles ? A->num_tiles : A->M;
    int W = ((by_tiles && A->num_long > 0) || path == kPathBinned) ? 1 : pick_windows(A->M, A->N, units);
    p->windows = W;
    p->unit.assign(W + 1, 0);
    p->row.assign(W + 1, 0);
    std::vector<long long> elem(W + 1, 0);
    for (int w = 0; w <= W; ++w) p->unit[w] = (int)((long long)units * w / W);
    for (int w = 0; w <= W; ++w) {
        if (by_tiles) {
            int2 t;
            SPMV_TRY_CUDA(cudaMemcpy(&t, A->tiles + p->unit[w], sizeof t, cudaMemcpyDeviceToHost));
            p->row[w] = t.x;
            elem[w] = t.y;
        } else {
            int off = 0;
            SPMV_TRY_CUDA(cudaMemcpy(&off, A->row_ptr + p->unit[w], sizeof off, cudaMemcpyDeviceToHost));
            p->row[w] = p->unit[w];
            elem[w] = off;
        }
    }
    SPMV_TRY(plan_needs(p, A->col_idx, elem));
    p->path = (int)path;
    return SPMV_B200_OK;
}

// ---- timing harness (reference protocol, main_cuda.cu:159-200) --------------------------------------------------
template <class Launch>
static int time_products(long long M, long long N, const double *x, double *y, double *d_x, double *d_y, int warmup, int iters,
                         double *mean_seconds, double *min_seconds, Launch launch) {
    if (iters <= 0 || warmup < 0) return fail(SPMV_B200_ERR_INVALID, "time: iters must be positive");
    cudaEvent_t a = nullptr, b = nullptr;
    auto body = [&]() -> int {
        SPMV_TRY_CUDA(cudaEventCreate(&a));
        SPMV_TRY_CUDA(cudaEventCreate(&b));
        if (N && x) SPMV_TRY_CUDA(cudaMemcpy(d_x, x, (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
        double sum = 0.0, best = 0.0;
        for (int i = 0; i < warmup + iters; ++i) {
            SPMV_TRY_CUDA(cudaEventRecord(a, nullptr));
            SPMV_TRY(launch());
            SPMV_TRY_CUDA(cudaEventRecord(b, nullptr));
            SPMV_TRY_CUDA(cudaEventSynchronize(b));
            float ms = 0.0f;
            SPMV_TRY_CUDA(cudaEventElapsedTime(&ms, a, b));
            if (i < warmup) continue;
            sum += ms * 1e-3;
            if (i == warmup || ms * 1e-3 < best) best = ms * 1e-3;
        }
        if (y && M) SPMV_TRY_CUDA(cudaMemcpy(y, d_y, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost));
        if (mean_seconds) *mean_seconds = sum / iters;
        if (min_seconds) *min_seconds = best;
        return SPMV_B200_OK;
    };
    const int rc = body();
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    return rc;
}

}  // namespace spmv

using namespace spmv;

extern "C" {

int spmv_b200_csr_spmv_host(spmv_b200_csr *A, const double *x, double *y, int accumulate, int algo) {
    if (!A || (A->M > 0 && !y) || (A->N > 0 && A->nnz > 0 && !x)) return fail(SPMV_B200_ERR_INVALID, "csr_spmv_host: NULL argument");
    if (algo < SPMV_B200_ALGO_AUTO || algo > SPMV_B200_ALGO_ROW) return fail(SPMV_B200_ERR_INVALID, "csr_spmv_host: unknown algo %d", algo);
    if (A->M == 0) return SPMV_B200_OK;
    if (!A->stage_x) SPMV_TRY_CUDA(cudaMalloc(&A->stage_x, std::max<size_t>(A->N, 1) * sizeof(double)));
    if (!A->stage_y) SPMV_TRY_CUDA(cudaMalloc(&A->stage_y, std::max<size_t>(A->M, 1) * sizeof(double)));
    if (!A->pipe) SPMV_TRY(pipe_create(&A->pipe));
    const CsrPath path = csr_resolve(A, algo);
    if (A->pipe->path != (int)path) SPMV_TRY(csr_plan_windows(A, path));
    HostPipe *p = A->pipe;
    const int rc = run_pipeline(p, A->M, x ? A->N : 0, x, y, A->stage_x, A->stage_y, accumulate != 0, [&](int w) {
        return csr_launch_window(A, path, p->unit[w], p->unit[w + 1], A->stage_x, A->stage_y, accumulate, p->compute);
    });
    if (rc != SPMV_B200_OK) abort_pipeline(p);
    return rc;
}

int spmv_b200_hll_spmv_host(spmv_b200_hll *H, const double *x, double *y) {
    if (!H || (H->M > 0 && !y) || (H->N > 0 && H->slots > 0 && !x)) return fail(SPMV_B200_ERR_INVALID, "hll_spmv_host: NULL argument");
    if (H->M == 0) return SPMV_B200_OK;
    if (!H->stage_x) SPMV_TRY_CUDA(cudaMalloc(&H->stage_x, std::max<size_t>(H->N, 1) * sizeof(double)));
    if (!H->stage_y) SPMV_TRY_CUDA(cudaMalloc(&H->stage_y, std::max<size_t>(H->M, 1) * sizeof(double)));
    if (!H->pipe) SPMV_TRY(pipe_create(&H->pipe));
    const HllPath path = hll_resolve(H);
    const bool stream_kernel = path == kHllStream;
    HostPipe *p = H->pipe;
    if (p->path != (int)path) {
        const int units = stream_kernel ? H->num_tiles : H->num_hacks;
        const int W = pick_windows(H->M, H->N, units);
        p->windows = W;
        p->unit.assign(W + 1, 0);
        p->row.assign(W + 1, 0);
        std::vector<long long> elem(W + 1, 0);
        for (int w = 0; w <= W; ++w) {
            p->unit[w] = (int)((long long)units * w / W);
            int hack = p->unit[w];
            if (stream_kernel) {
                HllTile t;
                SPMV_TRY_CUDA(cudaMemcpy(&t, H->tiles + p->unit[w], sizeof t, cudaMemcpyDeviceToHost));
                hack = t.hack;
            }
            p->row[w] = std::min<long long>((long long)hack * HACK_SIZE, H->M);
            elem[w] = H->host_off[hack];
        }
        SPMV_TRY(plan_needs(p, H->JA, elem));
        p->path = (int)path;
    }
    const int rc = run_pipeline(p, H->M, x ? H->N : 0, x, y, H->stage_x, H->stage_y, false, [&](int w) {
        return hll_launch_window(H, path, p->unit[w], p->unit[w + 1], H->stage_x, H->stage_y, p->compute);
    });
    if (rc != SPMV_B200_OK) abort_pipeline(p);
    return rc;
}

int spmv_b200_csr_time(spmv_b200_csr *A, const double *x, double *y, int algo, int warmup, int iters, double *mean_seconds,
                       double *min_seconds) {
    if (!A || (A->N > 0 && A->nnz > 0 && !x)) return fail(SPMV_B200_ERR_INVALID, "csr_time: NULL argument");
    if (!A->stage_x) SPMV_TRY_CUDA(cudaMalloc(&A->stage_x, std::max<size_t>(A->N, 1) * sizeof(double)));
    if (!A->stage_y) SPMV_TRY_CUDA(cudaMalloc(&A->stage_y, std::max<size_t>(A->M, 1) * sizeof(double)));
    return time_products(A->M, A->N, x, y, A->stage_x, A->stage_y, warmup, iters, mean_seconds, min_seconds,
                         [&]() { return spmv_b200_csr_spmv(A, A->stage_x, A->stage_y, 0, algo, nullptr); });
}

int spmv_b200_hll_time(spmv_b200_hll *H, const double *x, double *y, int kernel, int warmup, int iters, double *mean_seconds,
                       double *min_seconds) {
    if (!H || (H->N > 0 && H->slots > 0 && !x)) return fail(SPMV_B200_ERR_INVALID, "hll_time: NULL argument");
    if (kernel < 0 || kernel > 2) return fail(SPMV_B200_ERR_INVALID, "hll_time: kernel must be 0, 1 or 2");
    if (!H->stage_x) SPMV_TRY_CUDA(cudaMalloc(&H->stage_x, std::max<size_t>(H->N, 1) * sizeof(double)));
    if (!H->stage_y) SPMV_TRY_CUDA(cudaMalloc(&H->stage_y, std::max<size_t>(H->M, 1) * sizeof(double)));
    return time_products(H->M, H->N, x, y, H->stage_x, H->stage_y, warmup, iters, mean_seconds, min_seconds, [&]() {
        if (kernel == 1) return spmv_b200_hll_spmv_slice(H, H->stage_x, H->stage_y, nullptr);
        if (kernel == 2) return spmv_b200_hll_spmv_stream(H, H->stage_x, H->stage_y, nullptr);
        return spmv_b200_hll_spmv(H, H->stage_x, H->stage_y, nullptr);
    });
}

}  // extern "C"
