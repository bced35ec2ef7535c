// Micro-benchmark 5: can the bilinear parts of the pair loop move to the tensor pipe?
//   (a) issue rate of mma.sync.m16n8k8 tf32 (legacy warp-level path, HMMA) per SMSP
//   (b) the same with the loop's remaining per-output work interleaved: per 16x8 tile (4 outputs per thread)
//       4 MUFU.SQRT + 2 packed ops x k + FMNMX
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32 mma_tf32.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// NMMA mma per tile, MUFU: 4 sqrt per tile, PK: packed f32x2 ops per output pair
template <int NMMA, bool MUFU, int PK> __global__ void k(float *out, float seed, long long *cyc) {
    unsigned a[3][4], b[3][2];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 4; ++j) a[i][j] = __float_as_uint(seed + i + j) & 0xffffe000u;
        for (int j = 0; j < 2; ++j) b[i][j] = __float_as_uint(seed * 0.5f + i - j) & 0xffffe000u;
    }
    float best = 1e30f;
    const float2 kc = make_float2(seed, seed), wd = make_float2(1e-4f, 1e-4f);
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITER; ++it) {
        float c[4] = {seed, seed + 1.f, seed + 2.f, seed + 3.f};   // dd tile
        float q[4] = {0.f, 0.f, 0.f, 0.f};                         // q' tile
        b[0][0] += it;                                             // keep the loop body live
        if (NMMA >= 1) mma_tf32(c, a[0], b[0]);
        if (NMMA >= 2) mma_tf32(c, a[1], b[1]);
        if (NMMA >= 3) mma_tf32(q, a[2], b[2]);
        if (NMMA >= 4) mma_tf32(q, a[1], b[0]);
        if (NMMA >= 5) mma_tf32(c, a[2], b[1]);
        if (NMMA >= 6) mma_tf32(q, a[0], b[2]);
        float2 s0 = make_float2(c[0], c[1]), s1 = make_float2(c[2], c[3]);
        if (MUFU) { s0 = make_float2(sqrt_approx(c[0]), sqrt_approx(c[1])); s1 = make_float2(sqrt_approx(c[2]), sqrt_approx(c[3])); }
        float2 x0 = make_float2(q[0], q[1]), x1 = make_float2(q[2], q[3]);
#pragma unroll
        for (int p = 0; p < PK; ++p) { x0 = __ffma2_rn(s0, wd, x0); x1 = __ffma2_rn(s1, wd, x1); s0 = __fadd2_rn(s0, kc); s1 = __fadd2_rn(s1, kc); }
        best = fminf(best, fminf(fminf(x0.x, x0.y), fminf(x1.x, x1.y)));
        best = fminf(best, fminf(fminf(s0.x, s0.y), fminf(s1.x, s1.y)));   // keep both tiles live in every variant
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = best;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NMMA, bool MUFU, int PK> void run(const char *name) {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {8, 16, 32}) {
        k<NMMA, MUFU, PK><<<148, warps * 32>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
        k<NMMA, MUFU, PK><<<148, warps * 32>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-46s warps/SM=%2d  SMSP cycles per tile (4 outputs/thread) = %.2f\n", name, warps, avg / ((double)ITER * (warps / 4.0)));
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, false, 0>("1 mma");
    run<2, false, 0>("2 mma (dependent)");
    run<3, false, 0>("3 mma (2 dependent + 1)");
    run<6, false, 0>("6 mma (2 chains of 3)");
    run<0, true, 0>("4 MUFU.SQRT");
    run<0, true, 1>("4 MUFU + 2x(FFMA2+FADD2)");
    run<3, true, 0>("3 mma + 4 MUFU");
    run<3, true, 1>("3 mma + 4 MUFU + 2x(FFMA2+FADD2)");
    run<3, true, 2>("3 mma + 4 MUFU + 4x(FFMA2+FADD2)");
    run<4, true, 1>("4 mma + 4 MUFU + 2x(FFMA2+FADD2)");
    run<6, true, 1>("6 mma + 4 MUFU + 2x(FFMA2+FADD2)");
    return 0;
}
