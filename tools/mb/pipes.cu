// pipes.cu -- microbenchmark: issue cost of the bounds inner loop on sm_100a.
// Variants (per "pair-anchor" = one |x-y| max-accumulate + one x+y min-accumulate):
//   0: FADD, FMNMX(|.|), FADD, FMNMX            (4 instr / pair-anchor)
//   1: 4 FADD + 2 FMNMX3 per two anchors        (3 instr / pair-anchor)
//   2: 2 FADD2 + 2 FMNMX3 per two anchors       (2 instr / pair-anchor)
//   3: FADD only, 4: FMNMX only, 5: FMNMX3 only, 6: FADD2 only   (raw pipe rates)
#include <cstdio>
#include <cuda_runtime.h>
#define NACC 32
__device__ __forceinline__ float max3abs(float a, float b, float c)
{ float d; asm("max.abs.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float max3(float a, float b, float c)
{ float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c)
{ float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{ unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rc;
  asm("add.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb)); return *reinterpret_cast<float2 *>(&rc); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b)
{ unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rc;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb)); return *reinterpret_cast<float2 *>(&rc); }

template <int V> __global__ void __launch_bounds__(256, 1) k(float *out, const float *in, int iters)
{
    float lb[NACC], ub[NACC];
    float2 dj[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) { lb[q] = 0.f; ub[q] = 1e30f; dj[q] = make_float2(in[q * 2 + threadIdx.x], in[q * 2 + 1 + threadIdx.x]); }
    float2 di = make_float2(in[threadIdx.x + 100], in[threadIdx.x + 101]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            if (V == 0) {
                lb[q] = fmaxf(lb[q], fabsf(di.x - dj[q].x)); ub[q] = fminf(ub[q], di.x + dj[q].x);
                lb[q] = fmaxf(lb[q], fabsf(di.y - dj[q].y)); ub[q] = fminf(ub[q], di.y + dj[q].y);
            } else if (V == 1) {
                lb[q] = max3abs(lb[q], di.x - dj[q].x, di.y - dj[q].y);
                ub[q] = min3(ub[q], di.x + dj[q].x, di.y + dj[q].y);
            } else if (V == 2) {
                const float2 x = sub2(di, dj[q]), s = add2(di, dj[q]);
                lb[q] = max3abs(lb[q], x.x, x.y);
                ub[q] = min3(ub[q], s.x, s.y);
            } else if (V == 3) {
                lb[q] = lb[q] + di.x; ub[q] = ub[q] + di.y; lb[q] = lb[q] + dj[q].x; ub[q] = ub[q] + dj[q].y;
            } else if (V == 4) {
                lb[q] = fmaxf(lb[q], di.x); ub[q] = fminf(ub[q], di.y); lb[q] = fmaxf(lb[q], dj[q].x); ub[q] = fminf(ub[q], dj[q].y);
            } else if (V == 5) {
                lb[q] = max3(lb[q], di.x, dj[q].x); ub[q] = min3(ub[q], di.y, dj[q].y);
                lb[q] = max3(lb[q], di.y, dj[q].y); ub[q] = min3(ub[q], di.x, dj[q].x);
            } else if (V == 6) {
                float2 t = add2(make_float2(lb[q], ub[q]), dj[q]); t = add2(t, di); t = add2(t, dj[q]); t = add2(t, di);
                lb[q] = t.x; ub[q] = t.y;
            }
        }
        di.x += 1e-3f; di.y -= 1e-3f;   // keep the loop from being hoisted
    }
    float s = 0;
#pragma unroll
    for (int q = 0; q < NACC; ++q) s += lb[q] + ub[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V> void run(const char *name, int instr_per_q)
{
    float *out, *in; cudaMalloc(&out, 148 * 256 * 4 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    const int iters = 20000; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int nb = 1; nb <= 2; ++nb) {   // nb*8 warps per SM
        k<V><<<148 * nb, 256>>>(out, in, 100);
        cudaEventRecord(a); k<V><<<148 * nb, 256>>>(out, in, iters); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double winst = (double)iters * NACC * instr_per_q * 8.0 * nb;  // warp-instructions per SM
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        const double cyc = ms * 1e-3 * clk * 1e3;
        printf("%-28s warps/SM=%2d  %.3f ms  warp-instr/clk/SM = %.2f (of 4)  [nominal clk %d kHz]\n", name, 8 * nb, ms, winst / cyc, clk);
    }
}
int main()
{
    run<0>("v0 FADD+FMNMX x2 (8/q)", 8); run<1>("v1 4FADD+2FMNMX3 (6/q)", 6); run<2>("v2 2FADD2+2FMNMX3 (4/q)", 4);
    run<3>("FADD only (4/q)", 4); run<4>("FMNMX only (4/q)", 4); run<5>("FMNMX3 only (4/q)", 4); run<6>("FADD2 only (4/q)", 4);
    cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
    return 0;
}
