// common.cuh -- shared declarations for libannb.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/annb.h"

#define ANNB_API extern "C" __attribute__((visibility("default")))

namespace annb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr int kMaxAnchors = 64;

void set_error(const char *fmt, ...);
extern int64_t g_launches;
extern bool g_debug_sync;
extern bool g_trace;  // ANNB_TRACE=1: wall-clock of every index entry point and large allocation on stderr
double now_ms();
struct TraceScope {
    const char *name;
    double t0;
    explicit TraceScope(const char *n) : name(n), t0(g_trace ? now_ms() : 0.0) {}
    ~TraceScope()
    {
        if (g_trace) fprintf(stderr, "[annb-trace] %-34s %9.3f ms\n", name, now_ms() - t0);
    }
};

#define ANNB_CUDA(expr)                                                                  \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            annb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                            cudaGetErrorString(_e));                                     \
            return _e == cudaErrorMemoryAllocation ? ANNB_ENOMEM : ANNB_ECUDA;           \
        }                                                                                \
    } while (0)

#define ANNB_TRY(expr)              \
    do {                            \
        int _r = (expr);            \
        if (_r != ANNB_OK) return _r; \
    } while (0)

#define ANNB_REQUIRE(cond, code, ...)    \
    do {                                 \
        if (!(cond)) {                   \
            annb::set_error(__VA_ARGS__); \
            return (code);               \
        }                                \
    } while (0)

// every kernel launch goes through this so gpu_launches can be reported
#define ANNB_LAUNCH(kernel, grid, block, smem, stream, ...)                       \
    do {                                                                          \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);               \
        ++annb::g_launches;                                                       \
        ANNB_CUDA(cudaGetLastError());                                            \
        if (annb::g_debug_sync) { /* ANNB_DEBUG_SYNC=1: localise asynchronous faults */ \
            cudaError_t _s = cudaStreamSynchronize(stream);                       \
            if (_s != cudaSuccess) {                                              \
                annb::set_error("%s:%d: kernel %s failed: %s", __FILE__, __LINE__, #kernel, \
                                cudaGetErrorString(_s));                          \
                return ANNB_ECUDA;                                                \
            }                                                                     \
        }                                                                         \
    } while (0)

// Freed device blocks are kept in a per-process pool and handed out again (best fit) instead of
// going back to the driver: cudaMalloc / cudaFree of the multi-GB index buffers cost milliseconds
// each and cudaFree synchronises the device.  annb_pool_trim() / context destruction release them.
void *pool_take(size_t bytes, size_t *cap);
void pool_give(void *p, size_t cap);
void pool_trim();

// Device buffer with grow-only reuse (avoids cudaMalloc on the hot path).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return ANNB_OK;
        if (p) pool_give(p, cap);
        p = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        if ((p = pool_take(bytes, &cap)) != nullptr) return ANNB_OK;
        const double t0 = g_trace ? now_ms() : 0.0;
        cudaError_t e = cudaMalloc(&p, want);
        if (g_trace && want > (64u << 20))
            fprintf(stderr, "[annb-trace]   cudaMalloc %.1f MB %9.3f ms\n", want / 1048576.0, now_ms() - t0);
        if (e != cudaSuccess) {  // give cached blocks back to the driver and retry with the exact size
            cudaGetLastError();
            pool_trim();
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            p = nullptr;
            return ANNB_ENOMEM;
        }
        cap = want;
        return ANNB_OK;
    }
    void release()
    {
        if (p) pool_give(p, cap);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace annb

struct annb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // internal per-sweep timing
    cudaEvent_t uev0 = nullptr, uev1 = nullptr;  // annb_timer_start / stop
    int num_sms = annb::kNumSMs;
    size_t l2_bytes = 0;
    // scratch reused by the host-pointer entry points
    annb::DevBuf s_in[6], s_out[3];
    void *pinned = nullptr;
    size_t pinned_cap = 0;
};

enum annb_ds_kind { ANNB_DS_DENSE = 0, ANNB_DS_STRINGS = 1, ANNB_DS_HIST = 2 };

struct annb_dataset {
    annb_ctx *ctx = nullptr;
    int kind = 0;
    int dtype = 0;  // ANNB_F32 / ANNB_F64 for dense
    int64_t n = 0;
    int64_t d = 0;        // dense: dims; hist: bins; strings: max length
    int64_t ld = 0;       // dense/hist: row pitch in elements
    void *data = nullptr; // dense rows / hist CDFs (double) / string symbols (uint8, 16B-aligned rows)
    int64_t *offs = nullptr;  // strings: device start offset per string (n+1, 16B aligned starts)
    int32_t *lens = nullptr;  // strings: device length per string
    int sigma = 0;            // strings: alphabet size after remap
    int64_t max_len = 0;
    double *cost = nullptr;   // hist, general Wasserstein: (d, d) ground-cost matrix; `data` then holds unit masses
    size_t pooled_cap = 0;    // > 0: `data` came from the device-buffer pool (annb_dataset_gather) and returns to it
};

namespace annb {

// ---- internal device-pointer entry points shared between translation units ----
// All pointers are device pointers; work is enqueued on ctx->stream.

// out[p] = metric(X[i[p]], X[j[p]]), float64 out
int pair_dists_f64(annb_ctx *ctx, const annb_dataset *ds, int metric, const int32_t *i,
                   const int32_t *j, int64_t n, double *out);
// distances from one item (index read from device memory *anchor) to all n items
int anchor_row_f64(annb_ctx *ctx, const annb_dataset *ds, int metric, const int32_t *anchor_dev,
                   double *row);

int check_metric(const annb_dataset *ds, int metric);

}  // namespace annb
