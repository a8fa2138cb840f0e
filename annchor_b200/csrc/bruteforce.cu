// bruteforce.cu -- exact k-NN graph on the device: BruteForce.fit() (annchor/annchor.py:943-1023),
// the recall oracle at sizes where the CPU cannot compute it.
//
// Euclidean / cosine on dense rows (d <= 128) are a genuine dense contraction
//      ||x - y||^2 = x.x + y.y - 2 x.y^T,      cos(x, y) = 1 - x^.y^ (unit rows)
// and run on the 5th-generation tensor cores:
//   1. rows are split into two bf16 terms x = xh + xl (16 mantissa bits together) and laid out as
//        A' = [xh | xh | xl],  B' = [yh | yl | yh]   (K = 3 * d_pad)
//      so that ONE bf16 GEMM A' B'^T accumulates xh.yh + xh.yl + xl.yh in float32 -- the dropped
//      xl.yl term and the accumulation error are bounded by eps (bf_error_bound);
//   2. a persistent, warp-specialised kernel per 128-row block: one thread issues TMA loads
//      (cp.async.bulk.tensor, 128B swizzle) of the K-blocks, one thread issues tcgen05.mma into a
//      double-buffered TMEM accumulator (2 x 128 columns), four epilogue warps read the 128 x 128 tile
//      back with tcgen05.ld and keep, per row, the CAND smallest approximate values with their ids
//      (fused top-C: the N x N matrix never exists);
//   3. the CAND candidates of every row are re-ranked with the exact metric kernel (metrics.cu, the
//      same arithmetic as get_exact_ijs), and a row is accepted only if
//        exact k-th value + 2 eps < smallest approximate value that was NOT kept,
//      which proves that no pair outside the candidate list can be among the k nearest; rows that
//      fail (none on the bench data) are recomputed by the exact one-to-all kernel.
// So the result is the exact k-NN graph in the metric kernels' arithmetic; the tensor cores only
// prune.  Other metrics / wider rows: chunked all-pairs through the pair kernels.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <numeric>

#include "common.cuh"

namespace annb {
namespace bf {

constexpr int BM = 128, BN = 128, BK = 64;  // tile rows / columns / K elements (64 bf16 = one 128 B swizzle row)
constexpr int STAGES = 3;                   // B pipeline depth (16 KB per stage)
constexpr int MAXKB = 6;                    // K-blocks: 3 * d_pad / 64 with d_pad <= 128
constexpr int CAND = 64;                    // candidates kept per row
constexpr int TILE_BYTES = BM * BK * 2;     // 16 KB
constexpr int THREADS = 192;                // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, float32 accumulation; one thread issues for the CTA
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of float32 accumulators -> 32 registers per thread (thread = TMEM lane = tile row)
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor of a K-major bf16 tile [rows][64] written by TMA with the 128 B
// swizzle (cute::UMMA::SmemDescriptor): start address >> 4 | SBO = 1024 B (8 rows x 128 B) >> 4 at bit 32 |
// version 1 at bit 46 | layout SWIZZLE_128B (2) at bit 61; LBO is unused for swizzled K-major tiles
__device__ __forceinline__ uint64_t smem_desc_sw128(const void *tile, int k_elems)
{
    const uint32_t addr = smem_u32(tile) + (uint32_t)k_elems * 2u;  // advance inside the 128 B swizzle row
    uint64_t d = (uint64_t)((addr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;                    // leading byte offset (ignored), canonical value 1
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- 1. operand preparation -----------------------------------------------------------------------
// one warp per row: (optional) normalisation, bf16 hi / lo split, squared norm of what the GEMM sees
template <typename T>
__global__ void __launch_bounds__(256)
split_rows_kernel(const T *__restrict__ X, int64_t ld, int64_t n, int d, int dp, int64_t npad, int cosine,
                  __nv_bfloat16 *__restrict__ A, __nv_bfloat16 *__restrict__ B, float *__restrict__ norm)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= npad) return;
    const int K = 3 * dp;
    __nv_bfloat16 *a = A + row * K, *b = B + row * K;
    float scale = 1.0f;
    if (row < n && cosine) {
        double s = 0.0;
        for (int c = lane; c < d; c += 32) {
            const double v = (double)X[row * ld + c];
            s += v * v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        scale = s > 0.0 ? (float)(1.0 / sqrt(s)) : 0.0f;
    }
    float nn = 0.0f;
    for (int c = lane; c < dp; c += 32) {
        const float v = (row < n && c < d) ? (float)X[row * ld + c] * scale : 0.0f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        a[c] = h;
        a[dp + c] = h;
        a[2 * dp + c] = l;
        b[c] = h;
        b[dp + c] = l;
        b[2 * dp + c] = h;
        nn += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    if (lane == 0) norm[row] = row < n ? (cosine ? 0.5f : nn) : INFINITY;  // padding rows never rank
}

// ---- 2. GEMM + fused per-row top-CAND ---------------------------------------------------------------
struct GemmArgs {
    int64_t n, npad;
    int n_rb, n_ct, kb;  // row blocks, column tiles, K-blocks
    float scale;         // value = nx + ny - scale * (x . y): 2 (euclidean^2) or 1 (cosine, norms = 1/2)
    const float *norm;
    int32_t *cand_id;    // [npad][CAND]
    float *cand_val;     // [npad][CAND]
    float *cand_tau;     // [npad] largest kept value = lower bound of everything that was dropped
};

__global__ void __launch_bounds__(THREADS, 1)
gemm_topc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const GemmArgs G)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024 B alignment for the swizzled tiles
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = smem;                                   // [kb][128][64] bf16
    unsigned char *sB = sA + MAXKB * TILE_BYTES;                // [STAGES][128][64] bf16
    float *lv = reinterpret_cast<float *>(sB + STAGES * TILE_BYTES);  // [CAND][128] candidate values (transposed)
    int32_t *li = reinterpret_cast<int32_t *>(lv + CAND * BM);          // [CAND][128] candidate ids
    float *sNy = reinterpret_cast<float *>(li + CAND * BM);             // [2][128] column norms, per accumulator stage
    uint64_t *bars = reinterpret_cast<uint64_t *>(sNy + 2 * BN);
    uint64_t *full = bars, *empty = bars + STAGES, *a_full = bars + 2 * STAGES, *a_empty = a_full + 1;
    uint64_t *t_full = a_empty + 1, *t_empty = t_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&t_full[s], 1);
            mbar_init(&t_empty[s], 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 2 accumulators x 128 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t st = 0, ph = 0, pa = 0;
            for (int rb = blockIdx.x; rb < G.n_rb; rb += gridDim.x) {
                mbar_wait(a_empty, pa ^ 1);  // the MMAs of the previous row block have read the A tiles
                mbar_expect_tx(a_full, (uint32_t)(G.kb * TILE_BYTES));
                for (int k = 0; k < G.kb; ++k) tma_load_2d(sA + k * TILE_BYTES, &mapA, a_full, k * BK, rb * BM);
                pa ^= 1;
                for (int ct = 0; ct < G.n_ct; ++ct)
                    for (int k = 0; k < G.kb; ++k) {
                        mbar_wait(&empty[st], ph ^ 1);
                        mbar_expect_tx(&full[st], TILE_BYTES);
                        tma_load_2d(sB + st * TILE_BYTES, &mapB, &full[st], k * BK, ct * BN);
                        if (++st == STAGES) {
                            st = 0;
                            ph ^= 1;
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_bf16(BM, BN);
            uint32_t st = 0, ph = 0, pa = 0, acc = 0, pacc = 0;
            for (int rb = blockIdx.x; rb < G.n_rb; rb += gridDim.x) {
                mbar_wait(a_full, pa);
                pa ^= 1;
                for (int ct = 0; ct < G.n_ct; ++ct) {
                    mbar_wait(&t_empty[acc], pacc ^ 1);  // the epilogue has drained this accumulator
                    tc_fence_after();
                    for (int k = 0; k < G.kb; ++k) {
                        mbar_wait(&full[st], ph);
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            tc_mma_bf16(tmem_base + acc * BN, smem_desc_sw128(sA + k * TILE_BYTES, kk * 16),
                                        smem_desc_sw128(sB + st * TILE_BYTES, kk * 16), idesc, (k | kk) != 0);
                        tc_commit(&empty[st]);  // frees the B stage once these MMAs have completed
                        if (++st == STAGES) {
                            st = 0;
                            ph ^= 1;
                        }
                    }
                    tc_commit(&t_full[acc]);  // accumulator ready for the epilogue
                    if (++acc == 2) {
                        acc = 0;
                        pacc ^= 1;
                    }
                }
                tc_commit(a_empty);  // all MMAs reading this row block's A tiles have completed
            }
        }
    } else {
        // ===== epilogue: four warps, warp w reads TMEM lanes 32 * (w % 4) .. + 31 =====
        const int q = warp & 3;
        const int r = q * 32 + lane;      // tile row = TMEM lane
        const int et = threadIdx.x - 64;  // 0..127 among the epilogue threads
        uint32_t acc = 0, pacc = 0;
        for (int rb = blockIdx.x; rb < G.n_rb; rb += gridDim.x) {
            const int64_t gi = (int64_t)rb * BM + r;
            const float nx = G.norm[gi];
            for (int k = 0; k < CAND; ++k) {
                lv[k * BM + r] = INFINITY;
                li[k * BM + r] = -1;
            }
            float tau = INFINITY;
            int pmax = 0;
            for (int ct = 0; ct < G.n_ct; ++ct) {
                sNy[acc * BN + et] = G.norm[(int64_t)ct * BN + et];
                asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
                mbar_wait(&t_full[acc], pacc);
                tc_fence_after();
                const float *ny = sNy + acc * BN;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tc_ld_32x32(tmem_base + acc * BN + c0 + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float val = fmaf(-G.scale, __uint_as_float(v[c]), nx + ny[c0 + c]);
                        if (val < tau) {
                            const int64_t gj = (int64_t)ct * BN + c0 + c;
                            if (gj != gi) {
                                lv[pmax * BM + r] = val;
                                li[pmax * BM + r] = (int32_t)gj;
                                tau = -INFINITY;
                                for (int k = 0; k < CAND; ++k) {
                                    const float x = lv[k * BM + r];
                                    if (x > tau) {
                                        tau = x;
                                        pmax = k;
                                    }
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[acc]);
                if (++acc == 2) {
                    acc = 0;
                    pacc ^= 1;
                }
            }
            for (int k = 0; k < CAND; ++k) {
                G.cand_id[gi * CAND + k] = li[k * BM + r];
                G.cand_val[gi * CAND + k] = lv[k * BM + r];
            }
            G.cand_tau[gi] = tau;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

constexpr size_t GEMM_SMEM = 1024 + (size_t)(MAXKB + STAGES) * TILE_BYTES + (size_t)CAND * BM * 8 + 2 * BN * 4 +
                             (2 * STAGES + 6) * 8 + 64;

// cuTensorMapEncodeTiled through the runtime (libannb links cudart statically, not libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap *map, void *base, int64_t rows, int K)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        ANNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
        ANNB_REQUIRE(p != nullptr && qr == cudaDriverEntryPointSuccess, ANNB_ECUDA,
                     "cuTensorMapEncodeTiled is not available from this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ANNB_REQUIRE(r == CUDA_SUCCESS, ANNB_ECUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return ANNB_OK;
}

// pairs (row, candidate) of a block of rows for the exact re-rank; absent candidates pair the row with itself
__global__ void cand_pairs_kernel(const int32_t *__restrict__ cand_id, int64_t row0, int64_t rows,
                                  int32_t *__restrict__ I, int32_t *__restrict__ J)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < rows * CAND;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = row0 + p / CAND;
        const int32_t id = cand_id[row * CAND + p % CAND];
        I[p] = (int32_t)row;
        J[p] = id < 0 ? (int32_t)row : id;
    }
}
__global__ void all_pairs_kernel(int64_t row0, int64_t rows, int64_t n, int32_t *__restrict__ I, int32_t *__restrict__ J)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < rows * n;
         p += (int64_t)gridDim.x * blockDim.x) {
        I[p] = (int32_t)(row0 + p / n);
        J[p] = (int32_t)(p % n);
    }
}

}  // namespace bf
}  // namespace annb

using namespace annb;

namespace {

struct Cand {
    double d;
    int64_t id;
    bool operator<(const Cand &o) const { return d < o.d || (d == o.d && id < o.id); }
};

// k-1 nearest of a row from (ids, distances); column 0 = (row, 0)
void emit_row(int64_t row, std::vector<Cand> &c, int64_t k, int64_t *idx, double *dist)
{
    const size_t want = (size_t)std::min<int64_t>(k - 1, (int64_t)c.size());
    std::partial_sort(c.begin(), c.begin() + want, c.end());
    idx[row * k] = row;
    dist[row * k] = 0.0;
    for (int64_t q = 1; q < k; ++q) {
        const bool have = (size_t)(q - 1) < want;
        idx[row * k + q] = have ? c[q - 1].id : -1;
        dist[row * k + q] = have ? c[q - 1].d : INFINITY;
    }
}

// exact brute force of a set of rows through the pair kernels, in chunks of rows
int exact_rows(annb_ctx *c, const annb_dataset *ds, int metric, const std::vector<int64_t> &rows, bool contiguous,
               int64_t k, int64_t *idx, double *dist)
{
    const int64_t n = ds->n;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>((int64_t)rows.size(), (1 << 25) / n));
    DevBuf bi, bj, bd, brow;
    ANNB_TRY(bi.ensure((size_t)chunk * n * 4));
    ANNB_TRY(bj.ensure((size_t)chunk * n * 4));
    ANNB_TRY(bd.ensure((size_t)chunk * n * 8));
    std::vector<double> h((size_t)chunk * n);
    std::vector<int32_t> hi;
    int rc = ANNB_OK;
    for (size_t r0 = 0; r0 < rows.size() && rc == ANNB_OK; r0 += chunk) {
        const int64_t m = std::min<int64_t>(chunk, (int64_t)rows.size() - r0);
        if (contiguous) {
            bf::all_pairs_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(rows[r0], m, n, bi.as<int32_t>(), bj.as<int32_t>());
            ++g_launches;
        } else {
            hi.resize((size_t)m * n * 2);
            for (int64_t q = 0; q < m; ++q)
                for (int64_t j = 0; j < n; ++j) {
                    hi[(size_t)q * n + j] = (int32_t)rows[r0 + q];
                    hi[(size_t)m * n + q * n + j] = (int32_t)j;
                }
            cudaMemcpyAsync(bi.p, hi.data(), (size_t)m * n * 4, cudaMemcpyHostToDevice, c->stream);
            cudaMemcpyAsync(bj.p, hi.data() + (size_t)m * n, (size_t)m * n * 4, cudaMemcpyHostToDevice, c->stream);
        }
        rc = pair_dists_f64(c, ds, metric, bi.as<int32_t>(), bj.as<int32_t>(), m * n, bd.as<double>());
        if (rc != ANNB_OK) break;
        if (cudaMemcpyAsync(h.data(), bd.p, (size_t)m * n * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) {
            set_error("brute force: device error: %s", cudaGetErrorString(cudaGetLastError()));
            rc = ANNB_ECUDA;
            break;
        }
        std::vector<Cand> cand;
        for (int64_t q = 0; q < m; ++q) {
            const int64_t row = rows[r0 + q];
            cand.clear();
            for (int64_t j = 0; j < n; ++j)
                if (j != row) cand.push_back({h[(size_t)q * n + j], j});
            emit_row(row, cand, k, idx, dist);
        }
    }
    bi.release();
    bj.release();
    bd.release();
    return rc;
}

}  // namespace

// rigorous bound on |value the GEMM produced - value in exact arithmetic| for rows of squared norm <= m2:
// bf16 hi/lo split leaves |x - xh - xl| <= 2^-17 |x| per element and drops xl.yl (<= 2^-16 |x||y|);
// float32 accumulation of K products adds <= K 2^-23 |x||y|; the norm terms are float32 sums
static double bf_error_bound(double m2, int K, double scale)
{
    return scale * m2 * (std::ldexp(1.0, -15) + K * std::ldexp(1.0, -22)) + 4.0 * m2 * std::ldexp(1.0, -22);
}

ANNB_API int annb_bruteforce_knn(annb_ctx *c, const annb_dataset *ds, int metric, int64_t k, int64_t *idx,
                                 double *dist)
{
    TraceScope _ts("annb_bruteforce_knn");
    ANNB_REQUIRE(c && ds && idx && dist, ANNB_EINVAL, "NULL argument");
    ANNB_TRY(check_metric(ds, metric));
    const int64_t n = ds->n;
    ANNB_REQUIRE(k >= 1 && k <= n, ANNB_EINVAL, "k=%lld outside [1, n]", (long long)k);
    ANNB_CUDA(cudaSetDevice(c->device));
    const bool dense = ds->kind == ANNB_DS_DENSE && (metric == ANNB_EUCLIDEAN || metric == ANNB_COSINE);
    const bool force_exact = getenv("ANNB_BRUTEFORCE_EXACT") != nullptr;  // test knob: pair kernels only
    if (!dense || ds->d > 128 || n < 512 || k - 1 > bf::CAND / 2 || force_exact) {
        // all pairs through the metric kernels (strings, histograms, wide rows, tiny problems)
        std::vector<int64_t> rows(n);
        std::iota(rows.begin(), rows.end(), 0);
        return exact_rows(c, ds, metric, rows, true, k, idx, dist);
    }
    // ---- tensor-core path ----
    const int dp = ds->d <= 64 ? 64 : 128;
    const int K = 3 * dp;
    const int64_t npad = (n + bf::BM - 1) / bf::BM * bf::BM;
    DevBuf A, B, norm, cid, cval, ctau, bi, bj, bd;
    int rc = ANNB_OK;
    auto cleanup = [&]() {
        for (DevBuf *b : {&A, &B, &norm, &cid, &cval, &ctau, &bi, &bj, &bd}) b->release();
    };
#define BF_TRY(expr)           \
    do {                       \
        rc = (expr);           \
        if (rc != ANNB_OK) {   \
            cleanup();         \
            return rc;         \
        }                      \
    } while (0)
    BF_TRY(A.ensure((size_t)npad * K * 2));
    BF_TRY(B.ensure((size_t)npad * K * 2));
    BF_TRY(norm.ensure((size_t)npad * 4));
    BF_TRY(cid.ensure((size_t)npad * bf::CAND * 4));
    BF_TRY(cval.ensure((size_t)npad * bf::CAND * 4));
    BF_TRY(ctau.ensure((size_t)npad * 4));
    const int cosine = metric == ANNB_COSINE ? 1 : 0;
    const int sgrid = (int)((npad * 32 + 255) / 256);
    if (ds->dtype == ANNB_F32)
        bf::split_rows_kernel<float><<<sgrid, 256, 0, c->stream>>>((const float *)ds->data, ds->ld, n, (int)ds->d, dp,
                                                                  npad, cosine, A.as<__nv_bfloat16>(),
                                                                  B.as<__nv_bfloat16>(), norm.as<float>());
    else
        bf::split_rows_kernel<double><<<sgrid, 256, 0, c->stream>>>((const double *)ds->data, ds->ld, n, (int)ds->d,
                                                                   dp, npad, cosine, A.as<__nv_bfloat16>(),
                                                                   B.as<__nv_bfloat16>(), norm.as<float>());
    ++g_launches;
    CUtensorMap mapA, mapB;
    BF_TRY(bf::make_map(&mapA, A.p, npad, K));
    BF_TRY(bf::make_map(&mapB, B.p, npad, K));
    bf::GemmArgs G;
    G.n = n;
    G.npad = npad;
    G.n_rb = (int)(npad / bf::BM);
    G.n_ct = (int)(npad / bf::BN);
    G.kb = K / bf::BK;
    G.scale = cosine ? 1.0f : 2.0f;
    G.norm = norm.as<float>();
    G.cand_id = cid.as<int32_t>();
    G.cand_val = cval.as<float>();
    G.cand_tau = ctau.as<float>();
    if (cudaFuncSetAttribute(bf::gemm_topc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf::GEMM_SMEM) !=
        cudaSuccess) {
        set_error("brute force: cannot reserve %zu bytes of shared memory", bf::GEMM_SMEM);
        cleanup();
        return ANNB_ECUDA;
    }
    const int grid = std::min(c->num_sms, G.n_rb);
    bf::gemm_topc_kernel<<<grid, bf::THREADS, bf::GEMM_SMEM, c->stream>>>(mapA, mapB, G);
    ++g_launches;
    // largest squared norm -> error bound of the approximate values
    std::vector<float> hnorm(n), htau(n);
    if (cudaMemcpyAsync(hnorm.data(), norm.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaMemcpyAsync(htau.data(), ctau.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        set_error("brute force GEMM failed: %s", cudaGetErrorString(cudaGetLastError()));
        cleanup();
        return ANNB_ECUDA;
    }
    double m2 = 0.0;
    for (int64_t i = 0; i < n; ++i) m2 = std::max(m2, cosine ? 1.0 : (double)hnorm[i]);
    const double eps = bf_error_bound(m2, K, G.scale);
    // ---- exact re-rank of the candidates, certificate per row ----
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, (1 << 24) / bf::CAND));
    BF_TRY(bi.ensure((size_t)chunk * bf::CAND * 4));
    BF_TRY(bj.ensure((size_t)chunk * bf::CAND * 4));
    BF_TRY(bd.ensure((size_t)chunk * bf::CAND * 8));
    std::vector<double> hd((size_t)chunk * bf::CAND);
    std::vector<int32_t> hid((size_t)chunk * bf::CAND);
    std::vector<int64_t> redo;
    std::vector<Cand> cand;
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t m = std::min<int64_t>(chunk, n - r0);
        bf::cand_pairs_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(cid.as<int32_t>(), r0, m, bi.as<int32_t>(),
                                                                    bj.as<int32_t>());
        ++g_launches;
        BF_TRY(pair_dists_f64(c, ds, metric, bi.as<int32_t>(), bj.as<int32_t>(), m * bf::CAND, bd.as<double>()));
        if (cudaMemcpyAsync(hd.data(), bd.p, (size_t)m * bf::CAND * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaMemcpyAsync(hid.data(), cid.as<int32_t>() + r0 * bf::CAND, (size_t)m * bf::CAND * 4,
                            cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) {
            set_error("brute force re-rank failed: %s", cudaGetErrorString(cudaGetLastError()));
            cleanup();
            return ANNB_ECUDA;
        }
        for (int64_t q = 0; q < m; ++q) {
            const int64_t row = r0 + q;
            cand.clear();
            for (int s = 0; s < bf::CAND; ++s)
                if (hid[(size_t)q * bf::CAND + s] >= 0)
                    cand.push_back({hd[(size_t)q * bf::CAND + s], hid[(size_t)q * bf::CAND + s]});
            emit_row(row, cand, k, idx, dist);
            // certificate: everything that was dropped has approximate value >= tau, hence exact value
            // >= tau - eps; the k-th kept exact value (as the GEMM's quantity: d^2 or cosine) must be below
            if ((int64_t)cand.size() < n - 1) {
                const double dk = k > 1 ? dist[row * k + (k - 1)] : 0.0;
                const double ek = cosine ? dk : dk * dk;
                if (!(ek + 2.0 * eps < (double)htau[row])) redo.push_back(row);
            }
        }
    }
    if (g_trace)
        fprintf(stderr, "[annb-trace]   brute force: eps %.3g, %zu of %lld rows recomputed exactly\n", eps, redo.size(),
                (long long)n);
    if (!redo.empty()) BF_TRY(exact_rows(c, ds, metric, redo, false, k, idx, dist));
    cleanup();
#undef BF_TRY
    return ANNB_OK;
}
