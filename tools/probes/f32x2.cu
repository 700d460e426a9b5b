#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__global__ void k_add2(unsigned long long* out, int iters, unsigned long long b) {
    unsigned long long a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0+4, a5=a0+5, a6=a0+6, a7=a0+7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { a0 = add2(a0, b); a1 = add2(a1, b); a2 = add2(a2, b); a3 = add2(a3, b); a4 = add2(a4, b); a5 = add2(a5, b); a6 = add2(a6, b); a7 = add2(a7, b);}
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
__global__ void k_add1(float* out, int iters, float b) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0+4, a5=a0+5, a6=a0+6, a7=a0+7;
    float c0 = a0*2, c1=c0+1, c2=c0+2, c3=c0+3, c4=c0+4, c5=c0+5, c6=c0+6, c7=c0+7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { a0 += b; a1 += b; a2 += b; a3 += b; a4 += b; a5 += b; a6 += b; a7 += b;
                                      c0 += b; c1 += b; c2 += b; c3 += b; c4 += b; c5 += b; c6 += b; c7 += b; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + c0+c1+c2+c3+c4+c5+c6+c7;
}
int main() {
    void* out; cudaMalloc(&out, 148 * 8 * 512 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000; float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_add2<<<148 * 4, 512>>>((unsigned long long*)out, iters, 0x3f8000003f800000ull); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("add.f32x2: %.3f ms  %.2f Tadd/s (fp32 adds)\n", ms, 148.0 * 4 * 512 * iters * 64 * 2 / ms / 1e9);
        cudaEventRecord(e0); k_add1<<<148 * 4, 512>>>((float*)out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("add.f32  : %.3f ms  %.2f Tadd/s\n", ms, 148.0 * 4 * 512 * iters * 128 / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
