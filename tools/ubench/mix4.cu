// Micro-benchmark 5: two parents per thread -- one pair of LDS.128 feeds the packed arithmetic of TWO parent nodes
// (iteration = 26 FFMA2 + 2 FMUL + 6 MUFU + 2 LDS.128), vs one parent per thread (13 + 1 + 3 + 2 LDS.128).
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int NP> __global__ void k(float *out, float seed, long long *cyc) {
    __shared__ float4 tab[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tab[i] = make_float4(1.0001f, 0.9999f, 1.0002f, 0.9998f);
    __syncthreads();
    float2 a[NP][13]; float s[NP][4];
    for (int p = 0; p < NP; ++p) { for (int i = 0; i < 13; ++i) a[p][i] = make_float2(seed + i + p, seed - i); for (int i = 0; i < 4; ++i) s[p][i] = seed * i + 1.f + p; }
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < ITER; ++it) {
        const int b = (it * 2) & 510;
        const float4 t = tab[b], u = tab[b + 1];
        const float2 x0 = make_float2(t.x, t.y), x1 = make_float2(t.z, t.w), x2 = make_float2(u.x, u.y), x3 = make_float2(u.z, u.w);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            a[p][0] = __ffma2_rn(a[p][0], x0, c); a[p][1] = __ffma2_rn(a[p][1], x1, c);
            a[p][2] = __ffma2_rn(a[p][2], x2, c); a[p][3] = __ffma2_rn(a[p][3], x3, c);
            a[p][4] = __ffma2_rn(a[p][4], x0, c); a[p][5] = __ffma2_rn(a[p][5], x1, c);
#pragma unroll
            for (int i = 6; i < 13; ++i) a[p][i] = __ffma2_rn(a[p][i], m, c);
            s[p][3] = s[p][3] * m.x;
#pragma unroll
            for (int i = 0; i < 3; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[p][i]));
        }
    }
    long long t1 = clock64();
    float r = 0;
    for (int p = 0; p < NP; ++p) { for (int i = 0; i < 13; ++i) r += a[p][i].x + a[p][i].y; for (int i = 0; i < 4; ++i) r += s[p][i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NP> void run(const char *name) {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {8, 16, 32}) {
        k<NP><<<148, warps * 32>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
        k<NP><<<148, warps * 32>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-36s warps/SM=%2d  SMSP cycles per (parent, leaf pair) = %.2f\n", name, warps, avg / ((double)ITER * (warps / 4.0) * NP));
    }
}
int main() { run<1>("1 parent per thread, 2 LDS.128"); run<2>("2 parents per thread, 2 LDS.128"); run<4>("4 parents per thread, 2 LDS.128"); return 0; }
