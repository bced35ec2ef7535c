// Micro-benchmark 4: warp-uniform table delivered through KERNEL PARAMETERS (constant bank 0) so that it is
// fetched by the uniform datapath (LDCU -> UR operands of FFMA2) instead of LDS.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
struct Tab { float4 t[1024]; };
template <int MODE> __global__ void k(float *out, float seed, long long *cyc, const __grid_constant__ Tab tb) {
    float2 a[13]; float s[4];
    for (int i = 0; i < 13; ++i) a[i] = make_float2(seed + i, seed - i);
    for (int i = 0; i < 4; ++i) s[i] = seed * i + 1.f;
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        const int b = (it * 2) & 1022;
        const float4 t = tb.t[b], u = tb.t[b + 1];
        a[0] = __ffma2_rn(a[0], make_float2(t.x, t.y), c); a[1] = __ffma2_rn(a[1], make_float2(t.z, t.w), c);
        a[2] = __ffma2_rn(a[2], make_float2(u.x, u.y), c); a[3] = __ffma2_rn(a[3], make_float2(u.z, u.w), c);
        if (MODE == 1) {   // every table value used twice, as in the real loop (A and B each feed two FFMA2)
            a[4] = __ffma2_rn(a[4], make_float2(t.x, t.y), c); a[5] = __ffma2_rn(a[5], make_float2(t.z, t.w), c);
            a[6] = __ffma2_rn(a[6], make_float2(u.x, u.y), c); a[7] = __ffma2_rn(a[7], make_float2(u.z, u.w), c);
        } else {
            a[4] = __ffma2_rn(a[4], m, c); a[5] = __ffma2_rn(a[5], m, c); a[6] = __ffma2_rn(a[6], m, c); a[7] = __ffma2_rn(a[7], m, c);
        }
#pragma unroll
        for (int i = 8; i < 13; ++i) a[i] = __ffma2_rn(a[i], m, c);
        s[3] = s[3] * m.x;
#pragma unroll
        for (int i = 0; i < 3; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
    }
    long long t1 = clock64();
    float r = 0; for (int i = 0; i < 13; ++i) r += a[i].x + a[i].y; for (int i = 0; i < 4; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, const Tab &tb) {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {16, 32}) {
        k<MODE><<<148, warps * 32>>>(out, 1.5f, cyc, tb); cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, 1.5f, cyc, tb); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-52s warps/SM=%2d  SMSP cycles per iteration per warp = %.2f  (%s)\n", name, warps, avg / ((double)ITER * (warps / 4.0)), cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    static Tab tb; for (int i = 0; i < 1024; ++i) tb.t[i] = make_float4(1.0001f, 0.9999f, 1.0002f, 0.9998f);
    run<0>("base + 2 x 16B from kernel params (4 uses)", tb);
    run<1>("base + 2 x 16B from kernel params (8 uses)", tb);
    return 0;
}
