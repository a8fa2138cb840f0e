// anchors.cu -- K1: anchor pickers (annchor/pickers.py) and the get_exact_ijs host entry.
//
// MaxMinAnchorPicker.get_anchors (annchor/pickers.py:18-52) is n_anchors dependent rounds:
// distances anchor -> all, running min over rows >= 1, argmax -> next anchor.  All rounds are
// enqueued back to back on the context stream; the next anchor id never visits the host
// (the distance kernels read it from device memory).
#include "common.cuh"

namespace annb {

int pair_dists_f32_perm(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *I,
                        const int32_t *J, const int32_t *perm, int64_t n, float *out);

struct ArgMax {
    double v;
    int32_t i;
};

__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b)
{
    // np.argmax: first occurrence of the maximum
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

__device__ __forceinline__ ArgMax block_argmax(ArgMax x)
{
    __shared__ ArgMax sh[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        x = better(x, y);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        x = lane < nw ? sh[lane] : ArgMax{-INFINITY, INT32_MAX};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax y;
            y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
            y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
            x = better(x, y);
        }
    }
    return x;
}

// round 0: candidate value = row[j]; later rounds: minD[j] = min(minD[j], row[j])
// (anchor 0 never enters the min, annchor/pickers.py:47-50).
__global__ void __launch_bounds__(256)
maxmin_round_kernel(const double *__restrict__ row, double *__restrict__ minD, int64_t n,
                    int first_round, ArgMax *__restrict__ partial)
{
    ArgMax best{-INFINITY, INT32_MAX};
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x) {
        double v = row[j];
        if (!first_round) {
            v = fmin(minD[j], v);
            minD[j] = v;
        }
        best = better(best, ArgMax{v, (int32_t)j});
    }
    best = block_argmax(best);
    if (threadIdx.x == 0) partial[blockIdx.x] = best;
}

__global__ void __launch_bounds__(256)
maxmin_final_kernel(const ArgMax *__restrict__ partial, int nparts, int32_t *__restrict__ next,
                    int32_t *__restrict__ A_slot)
{
    ArgMax best{-INFINITY, INT32_MAX};
    for (int k = threadIdx.x; k < nparts; k += blockDim.x) best = better(best, partial[k]);
    best = block_argmax(best);
    if (threadIdx.x == 0) {
        *next = best.i;
        if (A_slot) *A_slot = best.i;
    }
}

__global__ void fill_f64_kernel(double *p, int64_t n, double v)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x)
        p[j] = v;
}

// (na, n) anchor-major -> (n, na) point-major
__global__ void transpose_D_kernel(const double *__restrict__ Dam, int64_t n, int na,
                                   double *__restrict__ Dpm)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * na;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t j = t / na;
        int a = (int)(t - j * na);
        Dpm[t] = Dam[(int64_t)a * n + j];
    }
}

__global__ void split_ij_kernel(const int64_t *__restrict__ ij, int64_t n, int32_t *__restrict__ I,
                                int32_t *__restrict__ J)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        I[p] = (int32_t)ij[2 * p];
        J[p] = (int32_t)ij[2 * p + 1];
    }
}

// A_dev: int32[na] (A_dev[0] must hold the first anchor), Dam: (na, n) float64 anchor-major.
int maxmin_device(annb_ctx *c, const annb_dataset *ds, int metric, int na, int32_t *A_dev,
                  double *Dam, DevBuf &scratch)
{
    const int64_t n = ds->n;
    const int parts = c->num_sms * 4;
    // scratch: minD[n] | partial[parts] | next
    size_t off_part = ((size_t)n * sizeof(double) + 255) / 256 * 256;
    size_t off_next = off_part + ((size_t)parts * sizeof(ArgMax) + 255) / 256 * 256;
    ANNB_TRY(scratch.ensure(off_next + 256));
    double *minD = scratch.as<double>();
    ArgMax *partial = reinterpret_cast<ArgMax *>(scratch.as<char>() + off_part);
    int32_t *next = reinterpret_cast<int32_t *>(scratch.as<char>() + off_next);
    ANNB_LAUNCH(fill_f64_kernel, parts, 256, 0, c->stream, minD, n, (double)INFINITY);
    ANNB_CUDA(cudaMemcpyAsync(next, A_dev, sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    for (int i = 0; i < na; ++i) {
        double *row = Dam + (int64_t)i * n;
        ANNB_TRY(anchor_row_f64(c, ds, metric, next, row));
        ANNB_LAUNCH(maxmin_round_kernel, parts, 256, 0, c->stream, row, minD, n, i == 0 ? 1 : 0,
                    partial);
        ANNB_LAUNCH(maxmin_final_kernel, 1, 256, 0, c->stream, partial, parts, next,
                    i + 1 < na ? A_dev + i + 1 : (int32_t *)nullptr);
    }
    return ANNB_OK;
}

int transpose_D(annb_ctx *c, const double *Dam, int64_t n, int na, double *Dpm)
{
    ANNB_LAUNCH(transpose_D_kernel, c->num_sms * 8, 256, 0, c->stream, Dam, n, na, Dpm);
    return ANNB_OK;
}

int split_ij(annb_ctx *c, const int64_t *ij_dev, int64_t n, int32_t *I, int32_t *J)
{
    if (n == 0) return ANNB_OK;
    ANNB_LAUNCH(split_ij_kernel, c->num_sms * 8, 256, 0, c->stream, ij_dev, n, I, J);
    return ANNB_OK;
}

}  // namespace annb

using namespace annb;

ANNB_API int annb_pair_dists(annb_ctx *c, const annb_dataset *ds, int metric, const int64_t *ij,
                             int64_t n, double *out)
{
    ANNB_REQUIRE(c && ds, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(n >= 0, ANNB_EINVAL, "n < 0");
    ANNB_TRY(check_metric(ds, metric));
    if (n == 0) return ANNB_OK;
    ANNB_REQUIRE(ij && out, ANNB_EINVAL, "NULL buffer");
    ANNB_CUDA(cudaSetDevice(c->device));
    for (int64_t p = 0; p < 2 * n; ++p)
        ANNB_REQUIRE(ij[p] >= 0 && ij[p] < ds->n, ANNB_EINVAL,
                     "pair %lld references item %lld outside [0,%lld)", (long long)(p / 2),
                     (long long)ij[p], (long long)ds->n);
    ANNB_TRY(c->s_in[0].ensure((size_t)n * 16));
    ANNB_TRY(c->s_in[1].ensure((size_t)n * 4));
    ANNB_TRY(c->s_in[2].ensure((size_t)n * 4));
    ANNB_TRY(c->s_out[0].ensure((size_t)n * 8));
    ANNB_CUDA(cudaMemcpyAsync(c->s_in[0].p, ij, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    ANNB_TRY(split_ij(c, c->s_in[0].as<int64_t>(), n, c->s_in[1].as<int32_t>(),
                      c->s_in[2].as<int32_t>()));
    ANNB_TRY(pair_dists_f64(c, ds, metric, c->s_in[1].as<int32_t>(), c->s_in[2].as<int32_t>(), n,
                            c->s_out[0].as<double>()));
    ANNB_CUDA(cudaMemcpyAsync(out, c->s_out[0].p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    return ANNB_OK;
}

ANNB_API int annb_pair_dists_dev(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *i,
                                 const int32_t *j, int64_t n, float *out)
{
    ANNB_REQUIRE(c && ds, ANNB_EINVAL, "NULL argument");
    ANNB_CUDA(cudaSetDevice(c->device));
    return pair_dists_f32_perm(c, ds, metric, i, j, nullptr, n, out);
}

ANNB_API int annb_maxmin_anchors(annb_ctx *c, const annb_dataset *ds, int metric, int64_t na,
                                 int64_t first, int64_t *A, double *D)
{
    ANNB_REQUIRE(c && ds, ANNB_EINVAL, "NULL argument");
    ANNB_TRY(check_metric(ds, metric));
    ANNB_REQUIRE(na > 0 && na <= ds->n, ANNB_EINVAL, "n_anchors=%lld out of range", (long long)na);
    ANNB_REQUIRE(first >= 0 && first < ds->n, ANNB_EINVAL, "first anchor out of range");
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t n = ds->n;
    ANNB_TRY(c->s_out[0].ensure((size_t)na * n * 8));  // anchor-major
    ANNB_TRY(c->s_out[1].ensure((size_t)na * n * 8));  // point-major
    ANNB_TRY(c->s_in[1].ensure((size_t)na * 4));
    int32_t f32 = (int32_t)first;
    ANNB_CUDA(cudaMemcpyAsync(c->s_in[1].p, &f32, 4, cudaMemcpyHostToDevice, c->stream));
    ANNB_TRY(maxmin_device(c, ds, metric, (int)na, c->s_in[1].as<int32_t>(),
                           c->s_out[0].as<double>(), c->s_in[0]));
    if (D) {
        ANNB_TRY(transpose_D(c, c->s_out[0].as<double>(), n, (int)na, c->s_out[1].as<double>()));
        ANNB_CUDA(cudaMemcpyAsync(D, c->s_out[1].p, (size_t)na * n * 8, cudaMemcpyDeviceToHost,
                                  c->stream));
    }
    std::vector<int32_t> a32(na);
    ANNB_CUDA(cudaMemcpyAsync(a32.data(), c->s_in[1].p, (size_t)na * 4, cudaMemcpyDeviceToHost,
                              c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    if (A)
        for (int64_t k = 0; k < na; ++k) A[k] = a32[k];
    return ANNB_OK;
}

ANNB_API int annb_anchor_dists(annb_ctx *c, const annb_dataset *ds, int metric, const int64_t *A,
                               int64_t na, double *D)
{
    ANNB_REQUIRE(c && ds && A && D, ANNB_EINVAL, "NULL argument");
    ANNB_TRY(check_metric(ds, metric));
    ANNB_REQUIRE(na > 0, ANNB_EINVAL, "n_anchors <= 0");
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t n = ds->n;
    std::vector<int32_t> a32(na);
    for (int64_t k = 0; k < na; ++k) {
        ANNB_REQUIRE(A[k] >= 0 && A[k] < n, ANNB_EINVAL, "anchor %lld out of range", (long long)A[k]);
        a32[k] = (int32_t)A[k];
    }
    ANNB_TRY(c->s_out[0].ensure((size_t)na * n * 8));
    ANNB_TRY(c->s_out[1].ensure((size_t)na * n * 8));
    ANNB_TRY(c->s_in[1].ensure((size_t)na * 4));
    ANNB_CUDA(cudaMemcpyAsync(c->s_in[1].p, a32.data(), (size_t)na * 4, cudaMemcpyHostToDevice,
                              c->stream));
    for (int64_t k = 0; k < na; ++k)
        ANNB_TRY(anchor_row_f64(c, ds, metric, c->s_in[1].as<int32_t>() + k,
                                c->s_out[0].as<double>() + k * n));
    ANNB_TRY(transpose_D(c, c->s_out[0].as<double>(), n, (int)na, c->s_out[1].as<double>()));
    ANNB_CUDA(cudaMemcpyAsync(D, c->s_out[1].p, (size_t)na * n * 8, cudaMemcpyDeviceToHost,
                              c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    return ANNB_OK;
}
