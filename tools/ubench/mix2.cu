// Micro-benchmark 3: cost of fetching warp-uniform table data inside the pair loop
// (base = 13 FFMA2 + 1 FMUL + 3 MUFU per iteration, independent chains).
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
__constant__ float4 ctab[512];
template <int MODE> __global__ void k(float *out, float seed, long long *cyc, const float4 *gtab) {
    __shared__ float4 tab[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tab[i] = make_float4(1.0001f, 0.9999f, 1.0002f, 0.9998f);
    __syncthreads();
    float2 a[13]; float s[4];
    for (int i = 0; i < 13; ++i) a[i] = make_float2(seed + i, seed - i);
    for (int i = 0; i < 4; ++i) s[i] = seed * i + 1.f;
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    const float2 *tab2 = reinterpret_cast<const float2 *>(tab);
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        float2 x0 = m, x1 = m, x2 = m, x3 = m, x4 = m, x5 = m, x6 = m, x7 = m;
        const int b = (it * 4) & 508;
        if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 7) { float4 t = tab[b]; x0 = make_float2(t.x, t.y); x1 = make_float2(t.z, t.w); }
        if (MODE == 2 || MODE == 3 || MODE == 7) { float4 t = tab[b + 1]; x2 = make_float2(t.x, t.y); x3 = make_float2(t.z, t.w); }
        if (MODE == 3) { float4 t = tab[b + 2], u = tab[b + 3]; x4 = make_float2(t.x, t.y); x5 = make_float2(t.z, t.w); x6 = make_float2(u.x, u.y); x7 = make_float2(u.z, u.w); }
        if (MODE == 4) { x0 = tab2[2 * b]; x1 = tab2[2 * b + 1]; x2 = tab2[2 * b + 2]; x3 = tab2[2 * b + 3]; }
        if (MODE == 5) { float4 t = __ldg(gtab + b), u = __ldg(gtab + b + 1); x0 = make_float2(t.x, t.y); x1 = make_float2(t.z, t.w); x2 = make_float2(u.x, u.y); x3 = make_float2(u.z, u.w); }
        if (MODE == 6) { float4 t = ctab[b], u = ctab[b + 1]; x0 = make_float2(t.x, t.y); x1 = make_float2(t.z, t.w); x2 = make_float2(u.x, u.y); x3 = make_float2(u.z, u.w); }
        a[0] = __ffma2_rn(a[0], x0, c); a[1] = __ffma2_rn(a[1], x1, c); a[2] = __ffma2_rn(a[2], x2, c); a[3] = __ffma2_rn(a[3], x3, c);
        a[4] = __ffma2_rn(a[4], x4, c); a[5] = __ffma2_rn(a[5], x5, c); a[6] = __ffma2_rn(a[6], x6, c); a[7] = __ffma2_rn(a[7], x7, c);
#pragma unroll
        for (int i = 8; i < 13; ++i) a[i] = __ffma2_rn(a[i], m, c);
        s[3] = s[3] * m.x;
        if (MODE != 7) {
#pragma unroll
            for (int i = 0; i < 3; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
        }
    }
    long long t1 = clock64();
    float r = 0; for (int i = 0; i < 13; ++i) r += a[i].x + a[i].y; for (int i = 0; i < 4; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, const float4 *g) {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {16, 32}) {
        k<MODE><<<148, warps * 32>>>(out, 1.5f, cyc, g); cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, 1.5f, cyc, g); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-40s warps/SM=%2d  SMSP cycles per iteration per warp = %.2f\n", name, warps, avg / ((double)ITER * (warps / 4.0)));
    }
}
int main() {
    float4 h[512]; for (int i = 0; i < 512; ++i) h[i] = make_float4(1.0001f, 0.9999f, 1.0002f, 0.9998f);
    float4 *g; cudaMalloc(&g, sizeof h); cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice); cudaMemcpyToSymbol(ctab, h, sizeof h);
    run<0>("base: 13 FFMA2 + FMUL + 3 MUFU", g);
    run<1>("base + 1 LDS.128", g);
    run<2>("base + 2 LDS.128", g);
    run<3>("base + 4 LDS.128", g);
    run<4>("base + 4 LDS.64", g);
    run<5>("base + 2 LDG.128 (uniform addr)", g);
    run<6>("base + 2 LDC.128 (__constant__)", g);
    run<7>("13 FFMA2 + FMUL + 2 LDS.128, no MUFU", g);
    return 0;
}
