// api_ctx.cu -- contexts, datasets, error plumbing of libannb.so.
#include <stdarg.h>

#include <time.h>
#include <mutex>
#include "common.cuh"

namespace annb {

static thread_local char g_err[1024] = "";
int64_t g_launches = 0;
bool g_debug_sync = getenv("ANNB_DEBUG_SYNC") != nullptr;
bool g_trace = getenv("ANNB_TRACE") != nullptr;
namespace {
struct PoolBlock {
    void *p;
    size_t cap;
    int dev;
};
std::vector<PoolBlock> g_pool;
std::mutex g_pool_mu;
}  // namespace

void *pool_take(size_t bytes, size_t *cap)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (int k = 0; k < (int)g_pool.size(); ++k) {
        const PoolBlock &b = g_pool[k];
        if (b.dev != dev || b.cap < bytes || b.cap > 2 * bytes + (8u << 20)) continue;
        if (best < 0 || b.cap < g_pool[best].cap) best = k;
    }
    if (best < 0) return nullptr;
    void *p = g_pool[best].p;
    *cap = g_pool[best].cap;
    g_pool.erase(g_pool.begin() + best);
    return p;
}

void pool_give(void *p, size_t cap)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool.size() >= 512) {  // keep the pool bounded: drop the smallest block
        int k0 = 0;
        for (int k = 1; k < (int)g_pool.size(); ++k)
            if (g_pool[k].cap < g_pool[k0].cap) k0 = k;
        cudaFree(g_pool[k0].p);
        g_pool.erase(g_pool.begin() + k0);
    }
    g_pool.push_back({p, cap, dev});
}

void pool_trim()
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (const PoolBlock &b : g_pool) {
        cudaSetDevice(b.dev);
        cudaFree(b.p);
    }
    cudaSetDevice(cur);
    g_pool.clear();
}

double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_metric(const annb_dataset *ds, int metric)
{
    ANNB_REQUIRE(ds != nullptr, ANNB_EINVAL, "dataset is NULL");
    switch (metric) {
    case ANNB_EUCLIDEAN:
    case ANNB_COSINE:
        ANNB_REQUIRE(ds->kind == ANNB_DS_DENSE, ANNB_EINVAL,
                     "euclidean/cosine need a dense dataset");
        break;
    case ANNB_LEVENSHTEIN:
        ANNB_REQUIRE(ds->kind == ANNB_DS_STRINGS, ANNB_EINVAL,
                     "levenshtein needs a strings dataset");
        break;
    case ANNB_WASSERSTEIN1D:
        ANNB_REQUIRE(ds->kind == ANNB_DS_HIST && ds->cost == nullptr, ANNB_EINVAL,
                     "wasserstein1d needs a histogram dataset (annb_dataset_hist)");
        break;
    case ANNB_WASSERSTEIN:
        ANNB_REQUIRE(ds->kind == ANNB_DS_HIST && ds->cost != nullptr, ANNB_EINVAL,
                     "wasserstein needs a histogram dataset with a cost matrix (annb_dataset_hist_cost)");
        break;
    default:
        set_error("unknown metric id %d", metric);
        return ANNB_EINVAL;
    }
    return ANNB_OK;
}

// histogram rows -> unit-mass CDFs (annchor/utils.py:82-84: kantorovich normalises each
// histogram to total mass 1).  One thread per row, sequential in bin order so the
// summation order is the textbook one.
template <typename T>
__global__ void hist_to_cdf_kernel(const T *__restrict__ H, int64_t n, int64_t nb,
                                   double *__restrict__ cdf)
{
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const T *h = H + r * nb;
    double s = 0.0;
    for (int64_t k = 0; k < nb; ++k) s += (double)h[k];
    double c = 0.0;
    double *o = cdf + r * nb;
    for (int64_t k = 0; k < nb; ++k) {
        c += (double)h[k] / s;
        o[k] = c;
    }
}

}  // namespace annb

using namespace annb;

ANNB_API const char *annb_last_error(void) { return g_err; }
ANNB_API int annb_version(void) { return 100; }
ANNB_API int64_t annb_launch_count(void) { return g_launches; }

ANNB_API int annb_ctx_create(int device, annb_ctx **out)
{
    ANNB_REQUIRE(out != nullptr, ANNB_EINVAL, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libannb has no CPU fallback",
                  cudaGetErrorString(e));
        return ANNB_ENOGPU;
    }
    ANNB_REQUIRE(device >= 0 && device < count, ANNB_EINVAL, "device %d out of range [0,%d)",
                 device, count);
    ANNB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ANNB_CUDA(cudaGetDeviceProperties(&prop, device));
    ANNB_REQUIRE(prop.major == 10, ANNB_ENOGPU,
                 "device %d is sm_%d%d; libannb is built for sm_100a (B200) only", device,
                 prop.major, prop.minor);
    annb_ctx *c = new annb_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    ANNB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    ANNB_CUDA(cudaEventCreate(&c->ev0));
    ANNB_CUDA(cudaEventCreate(&c->ev1));
    ANNB_CUDA(cudaEventCreate(&c->uev0));
    ANNB_CUDA(cudaEventCreate(&c->uev1));
    *out = c;
    return ANNB_OK;
}

ANNB_API int annb_pool_trim(void)
{
    pool_trim();
    return ANNB_OK;
}

ANNB_API int annb_ctx_destroy(annb_ctx *c)
{
    if (!c) return ANNB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &b : c->s_in) b.release();
    for (auto &b : c->s_out) b.release();
    pool_trim();
    if (c->pinned) cudaFreeHost(c->pinned);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->uev0);
    cudaEventDestroy(c->uev1);
    cudaStreamDestroy(c->stream);
    delete c;
    return ANNB_OK;
}

ANNB_API int annb_sync(annb_ctx *c)
{
    ANNB_REQUIRE(c != nullptr, ANNB_EINVAL, "ctx is NULL");
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    return ANNB_OK;
}

ANNB_API int annb_ctx_stream(annb_ctx *c, uint64_t *stream)
{
    ANNB_REQUIRE(c && stream, ANNB_EINVAL, "NULL argument");
    *stream = (uint64_t)(uintptr_t)c->stream;
    return ANNB_OK;
}

ANNB_API int annb_timer_start(annb_ctx *c)
{
    ANNB_REQUIRE(c != nullptr, ANNB_EINVAL, "ctx is NULL");
    ANNB_CUDA(cudaEventRecord(c->uev0, c->stream));
    return ANNB_OK;
}

ANNB_API int annb_timer_stop(annb_ctx *c, float *ms)
{
    ANNB_REQUIRE(c && ms, ANNB_EINVAL, "NULL argument");
    ANNB_CUDA(cudaEventRecord(c->uev1, c->stream));
    ANNB_CUDA(cudaEventSynchronize(c->uev1));
    ANNB_CUDA(cudaEventElapsedTime(ms, c->uev0, c->uev1));
    return ANNB_OK;
}

ANNB_API int annb_dataset_dense(annb_ctx *c, const void *X, int64_t n, int64_t d, int dtype,
                                int on_device, annb_dataset **out)
{
    ANNB_REQUIRE(c && X && out, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(n > 0 && d > 0, ANNB_EINVAL, "empty dataset (n=%lld, d=%lld)", (long long)n,
                 (long long)d);
    ANNB_REQUIRE(n < (1ll << 31), ANNB_ERANGE, "n=%lld exceeds int32 row ids", (long long)n);
    ANNB_REQUIRE(dtype == ANNB_F32 || dtype == ANNB_F64, ANNB_EINVAL,
                 "dense dtype must be ANNB_F32 or ANNB_F64");
    ANNB_CUDA(cudaSetDevice(c->device));
    size_t es = dtype == ANNB_F32 ? 4 : 8;
    // pad the row pitch to 16 bytes so rows can be read with 128-bit loads
    int64_t per16 = 16 / (int64_t)es;
    int64_t ld = (d + per16 - 1) / per16 * per16;
    annb_dataset *ds = new annb_dataset();
    ds->ctx = c;
    ds->kind = ANNB_DS_DENSE;
    ds->dtype = dtype;
    ds->n = n;
    ds->d = d;
    ds->ld = ld;
    cudaError_t e = cudaMalloc(&ds->data, (size_t)n * ld * es);
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete ds;
        set_error("cudaMalloc of dataset (%lld x %lld) failed: %s", (long long)n, (long long)d,
                  cudaGetErrorString(e));
        return ANNB_ENOMEM;
    }
    if (ld != d) ANNB_CUDA(cudaMemsetAsync(ds->data, 0, (size_t)n * ld * es, c->stream));
    ANNB_CUDA(cudaMemcpy2DAsync(ds->data, ld * es, X, d * es, d * es, n,
                                on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    *out = ds;
    return ANNB_OK;
}

ANNB_API int annb_dataset_strings(annb_ctx *c, const uint8_t *chars, const int64_t *offsets,
                                  int64_t n, annb_dataset **out)
{
    ANNB_REQUIRE(c && offsets && out, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(n > 0 && n < (1ll << 31), ANNB_EINVAL, "bad n=%lld", (long long)n);
    ANNB_CUDA(cudaSetDevice(c->device));
    // alphabet remap: byte value -> dense symbol id (keeps the Peq tables small)
    int map[256];
    for (int k = 0; k < 256; ++k) map[k] = -1;
    int sigma = 0;
    int64_t total = offsets[n], max_len = 0;
    for (int64_t t = 0; t < total; ++t)
        if (map[chars[t]] < 0) map[chars[t]] = 0;
    for (int k = 0; k < 256; ++k)
        if (map[k] == 0) map[k] = sigma++;
    if (sigma == 0) sigma = 1;
    std::vector<int64_t> offs(n + 1);
    std::vector<int32_t> lens(n);
    int64_t pos = 0;
    for (int64_t r = 0; r < n; ++r) {
        int64_t len = offsets[r + 1] - offsets[r];
        ANNB_REQUIRE(len >= 0, ANNB_EINVAL, "offsets not monotone at %lld", (long long)r);
        offs[r] = pos;
        lens[r] = (int32_t)len;
        if (len > max_len) max_len = len;
        pos += (len + 15) / 16 * 16;
    }
    offs[n] = pos;
    std::vector<uint8_t> packed((size_t)pos + 16, 0);
    for (int64_t r = 0; r < n; ++r) {
        const uint8_t *s = chars + offsets[r];
        uint8_t *o = packed.data() + offs[r];
        for (int32_t t = 0; t < lens[r]; ++t) o[t] = (uint8_t)map[s[t]];
    }
    annb_dataset *ds = new annb_dataset();
    ds->ctx = c;
    ds->kind = ANNB_DS_STRINGS;
    ds->dtype = ANNB_U8;
    ds->n = n;
    ds->d = max_len;
    ds->max_len = max_len;
    ds->sigma = sigma;
    ANNB_CUDA(cudaMalloc(&ds->data, packed.size()));
    ANNB_CUDA(cudaMalloc((void **)&ds->offs, (size_t)(n + 1) * sizeof(int64_t)));
    ANNB_CUDA(cudaMalloc((void **)&ds->lens, (size_t)n * sizeof(int32_t)));
    ANNB_CUDA(cudaMemcpy(ds->data, packed.data(), packed.size(), cudaMemcpyHostToDevice));
    ANNB_CUDA(cudaMemcpy(ds->offs, offs.data(), (size_t)(n + 1) * sizeof(int64_t),
                         cudaMemcpyHostToDevice));
    ANNB_CUDA(cudaMemcpy(ds->lens, lens.data(), (size_t)n * sizeof(int32_t),
                         cudaMemcpyHostToDevice));
    *out = ds;
    return ANNB_OK;
}

ANNB_API int annb_dataset_hist(annb_ctx *c, const void *H, int64_t n, int64_t nb, int dtype,
                               annb_dataset **out)
{
    ANNB_REQUIRE(c && H && out, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(n > 0 && nb > 0 && n < (1ll << 31), ANNB_EINVAL, "bad shape");
    ANNB_CUDA(cudaSetDevice(c->device));
    size_t es = dtype == ANNB_F32 ? 4 : dtype == ANNB_F64 ? 8 : dtype == ANNB_U8 ? 1 : 0;
    ANNB_REQUIRE(es != 0, ANNB_EINVAL, "hist dtype must be F32/F64/U8");
    annb_dataset *ds = new annb_dataset();
    ds->ctx = c;
    ds->kind = ANNB_DS_HIST;
    ds->dtype = ANNB_F64;
    ds->n = n;
    ds->d = nb;
    ds->ld = nb;
    ANNB_CUDA(cudaMalloc(&ds->data, (size_t)n * nb * sizeof(double)));
    ANNB_TRY(c->s_in[0].ensure((size_t)n * nb * es));
    ANNB_CUDA(cudaMemcpyAsync(c->s_in[0].p, H, (size_t)n * nb * es, cudaMemcpyHostToDevice,
                              c->stream));
    int grid = (int)((n + 127) / 128);
    if (dtype == ANNB_F32)
        ANNB_LAUNCH(hist_to_cdf_kernel<float>, grid, 128, 0, c->stream, c->s_in[0].as<float>(), n,
                    nb, (double *)ds->data);
    else if (dtype == ANNB_F64)
        ANNB_LAUNCH(hist_to_cdf_kernel<double>, grid, 128, 0, c->stream, c->s_in[0].as<double>(),
                    n, nb, (double *)ds->data);
    else
        ANNB_LAUNCH(hist_to_cdf_kernel<uint8_t>, grid, 128, 0, c->stream,
                    c->s_in[0].as<uint8_t>(), n, nb, (double *)ds->data);
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    *out = ds;
    return ANNB_OK;
}

namespace annb {
// histogram rows -> unit masses (annchor/utils.py:82-84), one thread per row, sums in bin order
template <typename T>
__global__ void hist_to_mass_kernel(const T *__restrict__ H, int64_t n, int64_t nb, double *__restrict__ mass)
{
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const T *h = H + r * nb;
    double s = 0.0;
    for (int64_t k = 0; k < nb; ++k) s += (double)h[k];
    for (int64_t k = 0; k < nb; ++k) mass[r * nb + k] = (double)h[k] / s;
}
}  // namespace annb

ANNB_API int annb_dataset_hist_cost(annb_ctx *c, const void *H, int64_t n, int64_t nb, int dtype,
                                    const double *cost, annb_dataset **out)
{
    ANNB_REQUIRE(c && H && cost && out, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(n > 0 && nb > 0 && n < (1ll << 31), ANNB_EINVAL, "bad shape");
    ANNB_REQUIRE(nb <= 64, ANNB_ERANGE, "the general-cost Wasserstein kernel supports at most 64 bins (got %lld)",
                 (long long)nb);
    for (int64_t q = 0; q < nb * nb; ++q)
        ANNB_REQUIRE(cost[q] >= 0.0 && cost[q] < INFINITY, ANNB_EINVAL, "cost matrix entries must be finite and >= 0");
    ANNB_CUDA(cudaSetDevice(c->device));
    size_t es = dtype == ANNB_F32 ? 4 : dtype == ANNB_F64 ? 8 : dtype == ANNB_U8 ? 1 : 0;
    ANNB_REQUIRE(es != 0, ANNB_EINVAL, "hist dtype must be F32/F64/U8");
    annb_dataset *ds = new annb_dataset();
    ds->ctx = c;
    ds->kind = ANNB_DS_HIST;
    ds->dtype = ANNB_F64;
    ds->n = n;
    ds->d = nb;
    ds->ld = nb;
    ANNB_CUDA(cudaMalloc(&ds->data, (size_t)n * nb * sizeof(double)));
    ANNB_CUDA(cudaMalloc((void **)&ds->cost, (size_t)nb * nb * sizeof(double)));
    ANNB_CUDA(cudaMemcpyAsync(ds->cost, cost, (size_t)nb * nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ANNB_TRY(c->s_in[0].ensure((size_t)n * nb * es));
    ANNB_CUDA(cudaMemcpyAsync(c->s_in[0].p, H, (size_t)n * nb * es, cudaMemcpyHostToDevice, c->stream));
    int grid = (int)((n + 127) / 128);
    if (dtype == ANNB_F32)
        ANNB_LAUNCH(hist_to_mass_kernel<float>, grid, 128, 0, c->stream, c->s_in[0].as<float>(), n, nb,
                    (double *)ds->data);
    else if (dtype == ANNB_F64)
        ANNB_LAUNCH(hist_to_mass_kernel<double>, grid, 128, 0, c->stream, c->s_in[0].as<double>(), n, nb,
                    (double *)ds->data);
    else
        ANNB_LAUNCH(hist_to_mass_kernel<uint8_t>, grid, 128, 0, c->stream, c->s_in[0].as<uint8_t>(), n, nb,
                    (double *)ds->data);
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    *out = ds;
    return ANNB_OK;
}

namespace annb {
// out row q = in row order[q]; rows are row_vec 8-byte words long, one warp per row
__global__ void gather_rows_kernel(const uint2 *__restrict__ src, int64_t row_vec, const int64_t *__restrict__ order,
                                   int64_t n, uint2 *__restrict__ dst)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = warp; q < n; q += nwarps) {
        const uint2 *s = src + order[q] * row_vec;
        uint2 *d = dst + q * row_vec;
        for (int64_t k = lane; k < row_vec; k += 32) d[k] = __ldg(s + k);
    }
}
// strings: 16-byte aligned starts on both sides
__global__ void gather_strings_kernel(const uint8_t *__restrict__ src, const int64_t *__restrict__ soffs,
                                      const int32_t *__restrict__ slens, const int64_t *__restrict__ order,
                                      const int64_t *__restrict__ doffs, int64_t n, uint8_t *__restrict__ dst,
                                      int32_t *__restrict__ dlens)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = warp; q < n; q += nwarps) {
        const int64_t o = order[q];
        const int32_t len = slens[o];
        const uint4 *s = reinterpret_cast<const uint4 *>(src + soffs[o]);
        uint4 *d = reinterpret_cast<uint4 *>(dst + doffs[q]);
        for (int k = lane; k < (len + 15) / 16; k += 32) d[k] = __ldg(s + k);
        if (lane == 0) dlens[q] = len;
    }
}
}  // namespace annb

// A new data set holding the items of `ds` in the order given (out item q = item order[q]); device to
// device.  Used to renumber the points of a fit in spatial order (annb_index_spatial_order).
ANNB_API int annb_dataset_gather(annb_ctx *c, const annb_dataset *ds, const int64_t *order, int64_t n,
                                 annb_dataset **out)
{
    ANNB_REQUIRE(c && ds && order && out, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(n > 0 && n < (1ll << 31), ANNB_EINVAL, "bad n");
    for (int64_t q = 0; q < n; ++q)
        ANNB_REQUIRE(order[q] >= 0 && order[q] < ds->n, ANNB_EINVAL, "gather index out of range");
    ANNB_CUDA(cudaSetDevice(c->device));
    ANNB_TRY(c->s_in[1].ensure((size_t)n * 8));
    ANNB_CUDA(cudaMemcpyAsync(c->s_in[1].p, order, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    annb_dataset *g = new annb_dataset(*ds);
    g->ctx = c;
    g->pooled_cap = 0;
    g->n = n;
    g->data = nullptr;
    g->offs = nullptr;
    g->lens = nullptr;
    g->cost = nullptr;
    if (ds->cost) {  // the copy owns its cost matrix
        ANNB_CUDA(cudaMalloc((void **)&g->cost, (size_t)ds->d * ds->d * 8));
        ANNB_CUDA(cudaMemcpyAsync(g->cost, ds->cost, (size_t)ds->d * ds->d * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    const int grid = c->num_sms * 8;
    if (ds->kind == ANNB_DS_STRINGS) {
        std::vector<int32_t> lens(ds->n);
        ANNB_CUDA(cudaMemcpy(lens.data(), ds->lens, (size_t)ds->n * 4, cudaMemcpyDeviceToHost));
        std::vector<int64_t> offs(n + 1);
        int64_t pos = 0;
        for (int64_t q = 0; q < n; ++q) {
            offs[q] = pos;
            pos += ((int64_t)lens[order[q]] + 15) / 16 * 16;
        }
        offs[n] = pos;
        ANNB_CUDA(cudaMalloc(&g->data, (size_t)pos + 16));
        ANNB_CUDA(cudaMalloc((void **)&g->offs, (size_t)(n + 1) * 8));
        ANNB_CUDA(cudaMalloc((void **)&g->lens, (size_t)n * 4));
        ANNB_CUDA(cudaMemcpyAsync(g->offs, offs.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        ANNB_LAUNCH(gather_strings_kernel, grid, 256, 0, c->stream, (const uint8_t *)ds->data, ds->offs, ds->lens,
                    c->s_in[1].as<int64_t>(), g->offs, n, (uint8_t *)g->data, g->lens);
        ANNB_CUDA(cudaStreamSynchronize(c->stream));  // offs (host vector) is read by the async copy
    } else {
        const size_t es = ds->dtype == ANNB_F32 ? 4 : 8;
        const size_t row_bytes = (size_t)ds->ld * es;
        ANNB_REQUIRE(row_bytes % 8 == 0, ANNB_ESTATE, "row pitch is not a multiple of 8 bytes");
        {   // through the pool: a fit() per step gathers the data set anew, cudaMalloc / cudaFree would cost ms each
            DevBuf b;
            ANNB_TRY(b.ensure((size_t)n * row_bytes));
            g->data = b.p;
            g->pooled_cap = b.cap;
        }
        ANNB_LAUNCH(gather_rows_kernel, grid, 256, 0, c->stream, (const uint2 *)ds->data, (int64_t)(row_bytes / 8),
                    c->s_in[1].as<int64_t>(), n, (uint2 *)g->data);
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
    }
    *out = g;
    return ANNB_OK;
}

ANNB_API int annb_dataset_free(annb_dataset *ds)
{
    if (!ds) return ANNB_OK;
    cudaSetDevice(ds->ctx->device);
    if (ds->data) {
        if (ds->pooled_cap) pool_give(ds->data, ds->pooled_cap);
        else cudaFree(ds->data);
    }
    if (ds->offs) cudaFree(ds->offs);
    if (ds->lens) cudaFree(ds->lens);
    if (ds->cost) cudaFree(ds->cost);
    delete ds;
    return ANNB_OK;
}

ANNB_API int64_t annb_dataset_len(const annb_dataset *ds) { return ds ? ds->n : -1; }
