// Micro-benchmark 2: the instruction MIX of the pass-1 pair loop with independent chains
// (13 packed FP + 1 scalar FMUL + 3 MUFU + 2 LDS.128 + 1 FMNMX3 per leaf pair), built up step by step.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int MODE> __global__ void k(float *out, float seed, long long *cyc) {
    __shared__ float4 tab[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tab[i] = make_float4(seed, seed + 1, seed + 2, seed + 3);
    __syncthreads();
    float2 a[13]; float s[4]; float best = 1e30f;
    for (int i = 0; i < 13; ++i) a[i] = make_float2(seed + i, seed - i);
    for (int i = 0; i < 4; ++i) s[i] = seed * i + 1.f;
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 13; ++i) a[i] = __ffma2_rn(a[i], m, c);
        s[3] = s[3] * m.x;
#pragma unroll
        for (int i = 0; i < 3; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
        if (MODE >= 1) {
            float4 t0_ = tab[(it * 2) & 511], t1_ = tab[(it * 2 + 1) & 511];
            a[0].x += t0_.x; a[1].y += t1_.w;      // 2 extra FADD to consume the loads (counted)
        }
        if (MODE >= 2) best = fminf(best, fminf(a[2].x, a[2].y));
    }
    long long t1 = clock64();
    float r = best; for (int i = 0; i < 13; ++i) r += a[i].x + a[i].y; for (int i = 0; i < 4; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name) {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {8, 16, 32}) {
        k<MODE><<<148, warps * 32>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-44s warps/SM=%2d  SMSP cycles per pair-iteration per warp = %.2f\n", name, warps, avg / ((double)ITER * (warps / 4.0)));
    }
}
int main() {
    run<0>("13 FFMA2 + FMUL + 3 MUFU");
    run<1>("13 FFMA2 + FMUL + 3 MUFU + 2 LDS.128 (+2 FADD)");
    run<2>("  ... + FMNMX3");
    return 0;
}
