// Microbenchmark + semantics probe for packed FP32 (f32x2) on sm_100a with contraction disabled.
//   nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3 -o f32x2_probe f32x2_probe.cu && ./f32x2_probe
// Questions: (1) does ptxas keep mul.rn.f32x2 + add.rn.f32x2 unfused?  (2) is an unfused packed mul/add pair (each written
// as fma.rn.f32x2 with a neutral operand) faster per flop than scalar FMUL + FADD when the kernel is issue bound?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2_raw(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2_raw(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__global__ void k_sem(const float* in, float* out) {
    float x = in[0], y = in[1], z = in[2];
    float s = x * y + z;                                   // scalar, --fmad=false
    u64 r = add2_raw(mul2_raw(pk(x, x), pk(y, y)), pk(z, z));
    float a, b; upk(r, a, b);
    u64 q = fma2(fma2(pk(x, x), pk(y, y), pk(-0.0f, -0.0f)), pk(1.0f, 1.0f), pk(z, z));
    float c, d; upk(q, c, d);
    out[0] = s; out[1] = a; out[2] = c; out[3] = fmaf(x, y, z);
}

template <int MODE>
__global__ void k_bench(float* out, int iters, float seed) {
    float acc[8]; u64 pacc[4];
    for (int k = 0; k < 8; ++k) acc[k] = seed + k;
    for (int k = 0; k < 4; ++k) pacc[k] = pk(seed + 2 * k, seed + 2 * k + 1);
    const float m = 1.0000001f, c = 1e-7f;
    const u64 m2 = pk(m, m), c2 = pk(c, c), one2 = pk(1.f, 1.f), nz2 = pk(-0.f, -0.f);
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = acc[k] * m + c;            // FMUL + FADD (fmad=false)
        } else if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) pacc[k] = fma2(fma2(pacc[k], m2, nz2), one2, c2);   // 2 x FFMA2, unfused semantics
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(acc[k], m, c);        // FFMA reference point
        }
    }
    float s = 0;
    for (int k = 0; k < 8; ++k) s += acc[k];
    for (int k = 0; k < 4; ++k) { float a, b; upk(pacc[k], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float h[3] = { 1.00000012f, 1.00000012f, -1.0f }, *d_in, *d_out, r[4];
    cudaMalloc(&d_in, 12); cudaMalloc(&d_out, 1 << 24);
    cudaMemcpy(d_in, h, 12, cudaMemcpyHostToDevice);
    k_sem<<<1, 1>>>(d_in, d_out); cudaMemcpy(r, d_out, 16, cudaMemcpyDeviceToHost);
    printf("semantics: scalar mul+add %.9g | mul.f32x2+add.f32x2 %.9g | fma2-neutral pair %.9g | true fma %.9g\n", r[0], r[1], r[2], r[3]);
    printf("  -> raw packed pair is %s ; neutral-fma pair is %s\n", r[1] == r[0] ? "UNFUSED (ok)" : "FUSED (ptxas contracted .rn ops!)", r[2] == r[0] ? "UNFUSED (ok)" : "FUSED");
    const int blocks = 148 * 8, threads = 256, iters = 1 << 14;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k_bench<0><<<blocks, threads>>>(d_out, iters, 1.0f);
            else if (mode == 1) k_bench<1><<<blocks, threads>>>(d_out, iters, 1.0f);
            else k_bench<2><<<blocks, threads>>>(d_out, iters, 1.0f);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double ops = (double)blocks * threads * iters * 8 * 2;   // mul + add per accumulator
        printf("mode %d (%s): %.3f ms  %.2f T(mul+add op)/s\n", mode, mode == 0 ? "scalar FMUL+FADD" : mode == 1 ? "packed 2xFFMA2 unfused" : "scalar FFMA fused", ms, ops / ms / 1e9);
    }
    return 0;
}
