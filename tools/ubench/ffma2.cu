// micro-benchmark: issue rate of scalar FFMA/FADD/FMUL against the packed f32x2 forms on sm_100a (8 independent chains per thread)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, int iters, float a, float b) {
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, int iters, float a, float b) {
    unsigned long long x[8], a2, b2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
    for (int i = 0; i < 8; ++i) { float lo = threadIdx.x + 2 * i, hi = lo + 1; asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(a2), "l"(b2));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * iters * 148 * 8 * 256;
        printf("scalar FFMA : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
        cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    return 0;
}
