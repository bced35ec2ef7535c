// Micro-benchmark: issue/pipe cost of packed f32x2 vs scalar FP32 vs MUFU on sm_100a.
// Each kernel runs ITER iterations of K independent dependency chains per thread; reports SMSP cycles
// per warp-instruction at full occupancy (clock64 deltas of one CTA per SM, 8..32 warps).
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE> __global__ void k(float *out, float seed, long long *cyc) {
    float2 a[8]; float s[8];
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed + i, seed - i); s[i] = seed * i + 1.f; }
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = __ffma2_rn(a[i], m, c);                               // 8 FFMA2
            if (MODE == 1) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }   // 16 FFMA
            if (MODE == 2) { a[i] = __ffma2_rn(a[i], m, c); s[i] = fmaf(s[i], m.x, c.x); }          // 8 FFMA2 + 8 FFMA
            if (MODE == 3) { a[i] = __ffma2_rn(a[i], m, c); if (i < 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[i])); }  // 8 FFMA2 + 3 MUFU
            if (MODE == 4) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); if (i < 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[i])); } // 16 FFMA + 3 MUFU
            if (MODE == 5) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[i])); }          // 8 MUFU
            if (MODE == 6) { a[i] = __ffma2_rn(a[i], m, c); s[i] = fminf(s[i], a[i].x); }           // 8 FFMA2 + 8 FMNMX (alu)
        }
    }
    long long t1 = clock64();
    float r = 0; for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, int instr_per_iter) {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {4, 8, 16, 32}) {
        k<MODE><<<148, warps * 32>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        double per = avg / ((double)ITER * instr_per_iter * (warps / 4.0));   // SMSP cycles per warp-instruction
        printf("%-28s warps/SM=%2d  cycles/warp-instr/SMSP = %.3f  (instr/iter=%d)\n", name, warps, per, instr_per_iter);
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FFMA2 x8", 8);
    run<1>("FFMA x16", 16);
    run<2>("FFMA2 x8 + FFMA x8", 16);
    run<3>("FFMA2 x8 + MUFU x3", 11);
    run<4>("FFMA x16 + MUFU x3", 19);
    run<5>("MUFU x8", 8);
    run<6>("FFMA2 x8 + FMNMX x8", 16);
    return 0;
}
