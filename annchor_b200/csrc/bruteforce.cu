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
//      (fused top-C: the N x N matrix never exists; a value costs one FMA and one compare against the
//      row's cut-off, the rare survivors are appended to a shared-memory buffer that the warp prunes
//      cooperatively with a bitonic network whenever it fills);
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
constexpr int STAGES = 2;                   // B pipeline depth (16 KB per stage; the tiles come from L2)
constexpr int MAXKB = 6;                    // K-blocks: 3 * d_pad / 64 with d_pad <= 128
constexpr int CAND = 64;                    // candidates kept per row
constexpr int CAP = 92;                     // per-row survivor buffer: CAND kept + room for new survivors
constexpr int CPITCH = CAP + 1;             // odd pitch: appends by 32 rows and reads along a row are conflict free
constexpr int CHECK = 8;                    // columns between two overflow checks (a row gains <= CHECK entries)
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
    int debug;           // ANNB_BF_DEBUG (timing experiments only): 1 = the epilogue drains TMEM without filtering,
                         // 2 = filters but never keeps, 3 = no MMAs (TMA pipeline alone)
    const float *norm;
    int32_t *cand_id;    // [npad][CAND]
    float *cand_val;     // [npad][CAND]
    float *cand_tau;     // [npad] largest kept value = lower bound of everything that was dropped
};

// compare-exchange with the lane `stride` away: the lower lane of the pair keeps the smaller element when `asc`
__device__ __forceinline__ void cx_shfl(float &v, int32_t &id, int stride, bool asc, int lane)
{
    const float pv = __shfl_xor_sync(0xffffffffu, v, stride);
    const int32_t pid = __shfl_xor_sync(0xffffffffu, id, stride);
    const bool keep_min = ((lane & stride) == 0) == asc;
    const int take = ((int)(pv < v) & (int)keep_min) | ((int)(pv > v) & (int)!keep_min);  // no divergent branch
    v = take ? pv : v;
    id = take ? pid : id;
}
__device__ __forceinline__ void cx_reg(float &a, int32_t &ia, float &b, int32_t &ib)  // a <= b afterwards
{
    const bool sw = b < a;
    const float lo = sw ? b : a, hi = sw ? a : b;
    const int32_t ilo = sw ? ib : ia, ihi = sw ? ia : ib;
    a = lo;
    b = hi;
    ia = ilo;
    ib = ihi;
}
// Warp-cooperative: keep the CAND (= 64) smallest of a row's cnt (<= CAP) buffered entries in slots 0..63 (any order;
// absent entries +inf / -1) and return the largest kept value.  128 slots as 4 registers per lane (slot = 32 k + lane):
// the two halves are sorted ascending by a bitonic network, then min(A[i], B[63 - i]) are the 64 smallest of all.
__device__ __forceinline__ float prune_row(float *rv, int32_t *ri, int cnt, int lane)
{
    float v[4];
    int32_t id[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = k * 32 + lane;
        v[k] = e < cnt ? rv[e] : INFINITY;
        id[k] = e < cnt ? ri[e] : -1;
    }
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride >= 1; stride >>= 1) {
            const bool asc = (lane & size) == 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) cx_shfl(v[k], id[k], stride, asc, lane);
        }
    // registers 1 and 3 descending, so that (0, 1) and (2, 3) are bitonic sequences of 64
#pragma unroll
    for (int k = 1; k < 4; k += 2) {
        v[k] = __shfl_sync(0xffffffffu, v[k], 31 - lane);
        id[k] = __shfl_sync(0xffffffffu, id[k], 31 - lane);
    }
    cx_reg(v[0], id[0], v[1], id[1]);
    cx_reg(v[2], id[2], v[3], id[3]);
#pragma unroll
    for (int stride = 16; stride >= 1; stride >>= 1)
#pragma unroll
        for (int k = 0; k < 4; ++k) cx_shfl(v[k], id[k], stride, true, lane);
    // halves sorted ascending: A = (v0, v1), B = (v2, v3); B[63 - i] sits in the other register at lane 31 - lane
    const float b1 = __shfl_sync(0xffffffffu, v[3], 31 - lane), b0 = __shfl_sync(0xffffffffu, v[2], 31 - lane);
    const int32_t j1 = __shfl_sync(0xffffffffu, id[3], 31 - lane), j0 = __shfl_sync(0xffffffffu, id[2], 31 - lane);
    const bool t0 = b1 < v[0], t1 = b0 < v[1];
    v[0] = t0 ? b1 : v[0];
    id[0] = t0 ? j1 : id[0];
    v[1] = t1 ? b0 : v[1];
    id[1] = t1 ? j0 : id[1];
    rv[lane] = v[0];
    ri[lane] = id[0];
    rv[32 + lane] = v[1];
    ri[32 + lane] = id[1];
    float t = fmaxf(v[0], v[1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    return t;
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_topc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const GemmArgs G)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024 B alignment for the swizzled tiles
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = smem;                                   // [kb][128][64] bf16
    unsigned char *sB = sA + MAXKB * TILE_BYTES;                // [STAGES][128][64] bf16
    float *cv = reinterpret_cast<float *>(sB + STAGES * TILE_BYTES);  // [128][CPITCH] survivor values per row
    int32_t *ci = reinterpret_cast<int32_t *>(cv + BM * CPITCH);       // [128][CPITCH] survivor ids
    float *sNy = reinterpret_cast<float *>(ci + BM * CPITCH);          // [4 warps][2 stages][128] column norms
    uint64_t *bars = reinterpret_cast<uint64_t *>(sNy + 8 * BN);
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
                            if (G.debug != 3)
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
        // A thread owns one row.  Values below the row's cut-off tau are APPENDED to the row's buffer (rare after the
        // first tiles); when a row has fewer than CHECK free slots the warp prunes it cooperatively to the CAND
        // smallest entries (bitonic sort in registers) and lowers tau to the largest of them.  tau only ever
        // decreases, so everything dropped -- at the filter or in a prune -- is >= the final tau.
        const int q = warp & 3;
        const int r = q * 32 + lane;      // tile row = TMEM lane
        uint32_t acc = 0, pacc = 0;
        for (int rb = blockIdx.x; rb < G.n_rb; rb += gridDim.x) {
            const int64_t gi = (int64_t)rb * BM + r;
            const float nx = G.norm[gi];
            float tau = G.debug == 2 ? -INFINITY : INFINITY;
            int cnt = 0;
            for (int ct = 0; ct < G.n_ct; ++ct) {
                // column norms of this tile: a private copy per warp and accumulator stage (no block barrier)
                float *ny = sNy + (q * 2 + acc) * BN;
                __syncwarp();
                *reinterpret_cast<float4 *>(ny + lane * 4) =
                    __ldg(reinterpret_cast<const float4 *>(G.norm + (int64_t)ct * BN) + lane);
                __syncwarp();
                mbar_wait(&t_full[acc], pacc);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tc_ld_32x32(tmem_base + acc * BN + c0 + ((uint32_t)(q * 32) << 16), v);
                    if (G.debug == 1) continue;
                    // all 32 values first (independent FMAs, the column norms come as 8 x LDS.128), one compare mask;
                    // the shared-memory appends happen only for the rare chunk that has a survivor
                    float val[32];
                    uint32_t mask = 0;
#pragma unroll
                    for (int c4 = 0; c4 < 32; c4 += 4) {
                        const float4 y4 = *reinterpret_cast<const float4 *>(ny + c0 + c4);
                        val[c4 + 0] = fmaf(-G.scale, __uint_as_float(v[c4 + 0]), nx + y4.x);
                        val[c4 + 1] = fmaf(-G.scale, __uint_as_float(v[c4 + 1]), nx + y4.y);
                        val[c4 + 2] = fmaf(-G.scale, __uint_as_float(v[c4 + 2]), nx + y4.z);
                        val[c4 + 3] = fmaf(-G.scale, __uint_as_float(v[c4 + 3]), nx + y4.w);
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) mask |= (val[c] < tau) ? (1u << c) : 0u;
                    if (!__any_sync(0xffffffffu, mask != 0u)) continue;
#pragma unroll
                    for (int g = 0; g < 32; g += CHECK) {
                        if ((mask >> g) & ((1u << CHECK) - 1u)) {
#pragma unroll
                            for (int c = g; c < g + CHECK; ++c)
                                if (((mask >> c) & 1u) && val[c] < tau) {  // tau may have dropped in a prune since the mask
                                    const int64_t gj = (int64_t)ct * BN + c0 + c;
                                    if (gj != gi) {
                                        cv[r * CPITCH + cnt] = val[c];
                                        ci[r * CPITCH + cnt] = (int32_t)gj;
                                        ++cnt;
                                    }
                                }
                        }
                        unsigned need = __ballot_sync(0xffffffffu, cnt > CAP - CHECK);
                        while (need) {
                            const int src = __ffs(need) - 1;
                            need &= need - 1;
                            const int c_r = __shfl_sync(0xffffffffu, cnt, src);
                            __syncwarp();
                            const float t = prune_row(cv + (q * 32 + src) * CPITCH, ci + (q * 32 + src) * CPITCH, c_r, lane);
                            if (lane == src) {
                                cnt = CAND;
                                tau = t;
                            }
                            __syncwarp();
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
            // final prune of every row, then the CAND slots go out (absent entries: id -1, value +inf)
            for (int src = 0; src < 32; ++src) {
                const int c_r = __shfl_sync(0xffffffffu, cnt, src);
                __syncwarp();
                const float t = prune_row(cv + (q * 32 + src) * CPITCH, ci + (q * 32 + src) * CPITCH, c_r, lane);
                if (lane == src) tau = t;
                __syncwarp();
                const int64_t go = (int64_t)rb * BM + q * 32 + src;
                for (int k = lane; k < CAND; k += 32) {
                    G.cand_id[go * CAND + k] = ci[(q * 32 + src) * CPITCH + k];
                    G.cand_val[go * CAND + k] = cv[(q * 32 + src) * CPITCH + k];
                }
            }
            G.cand_tau[gi] = tau;
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

constexpr size_t GEMM_SMEM = 1024 + (size_t)(MAXKB + STAGES) * TILE_BYTES + (size_t)BM * CPITCH * 8 + 8 * BN * 4 +
                             (2 * STAGES + 6) * 8 + 64;
static_assert(GEMM_SMEM <= 227 * 1024, "gemm_topc_kernel: shared memory budget");

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
// labels != nullptr: nearest ENEMIES -- only columns with another label take part, k columns, no self column
int exact_rows(annb_ctx *c, const annb_dataset *ds, int metric, const std::vector<int64_t> &rows, bool contiguous,
               int64_t k, int64_t *idx, double *dist, const int32_t *labels = nullptr)
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
            if (labels) {
                for (int64_t j = 0; j < n; ++j)
                    if (labels[j] != labels[row]) cand.push_back({h[(size_t)q * n + j], j});
                const size_t want = (size_t)std::min<int64_t>(k, (int64_t)cand.size());
                std::partial_sort(cand.begin(), cand.begin() + want, cand.end());
                for (int64_t t = 0; t < k; ++t) {
                    idx[row * k + t] = (size_t)t < want ? cand[t].id : -1;
                    dist[row * k + t] = (size_t)t < want ? cand[t].d : INFINITY;
                }
                continue;
            }
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

// Exact nearest-enemy graph (annchor/annchor.py:685-786 computes an approximation of it from the fitted state): for
// every item the nn nearest items carrying a DIFFERENT label, by exhaustive evaluation with the metric kernels.
ANNB_API int annb_nearest_enemies(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *labels, int64_t nn,
                                  int64_t *idx, double *dist)
{
    TraceScope _ts("annb_nearest_enemies");
    ANNB_REQUIRE(c && ds && labels && idx && dist, ANNB_EINVAL, "NULL argument");
    ANNB_TRY(check_metric(ds, metric));
    ANNB_REQUIRE(nn >= 1 && nn < ds->n, ANNB_EINVAL, "nn=%lld outside [1, n)", (long long)nn);
    ANNB_CUDA(cudaSetDevice(c->device));
    std::vector<int64_t> rows(ds->n);
    std::iota(rows.begin(), rows.end(), 0);
    return exact_rows(c, ds, metric, rows, true, nn, idx, dist, labels);
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
    G.debug = getenv("ANNB_BF_DEBUG") ? atoi(getenv("ANNB_BF_DEBUG")) : 0;
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
    if (g_trace) cudaEventRecord(c->ev0, c->stream);
    bf::gemm_topc_kernel<<<grid, bf::THREADS, bf::GEMM_SMEM, c->stream>>>(mapA, mapB, G);
    ++g_launches;
    if (g_trace) {
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        const double flop = 2.0 * (double)npad * (double)npad * (double)K;
        fprintf(stderr, "[annb-trace]   brute force gemm_topc_kernel: %.3f ms, %.1f TFLOP/s (bf16, K = %d), debug %d\n", ms,
                flop / (ms * 1e-3) / 1e12, K, G.debug);
    }
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
